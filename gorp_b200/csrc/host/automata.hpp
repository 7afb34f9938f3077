// DFA half of the definition compiler (load-time, host only).
//
//  * brics_min_dfa(): the language of `new RegExp(s, RegExp.NONE).toAutomaton(); minimize()` of
//    dk.brics.automaton 1.11-8 (call sites: reference autom/PolyMatcher.java:76-77) as a minimal, trimmed
//    DFA with maximal interval transitions.
//  * build_product(): the table that reference autom/Automata.java:57-124 builds (BFS product over the union
//    of interval start points, ids in discovery order, accept lists ascending) — bit-identical layout to
//    Automata._alphabet/_transitions/_accept, i.e. exactly what the Java-side DfaExport serialises.
//  * compact_tables(): results-invisible re-encoding for the GPU (merge identical columns, merge states that
//    are indistinguishable w.r.t. the FIRST accepting index — the only thing Gorp.extract reads, Gorp.java:166).
#pragma once
#include "common.hpp"

namespace gorp {

struct Interval { uint32_t lo, hi, to; };

struct MinDfa {
    std::vector<std::vector<Interval>> trans;  // per state, sorted by lo, maximal per destination
    std::vector<uint8_t> accept;
    int step(int s, uint32_t c) const;         // -1 == no transition (State.step returns null)
    std::vector<uint32_t> start_points() const;
};

MinDfa brics_min_dfa(const ustring& regex);  // throws std::invalid_argument with the brics-style message

struct DfaTables {                 // reference layout (what DfaExport ships)
    uint32_t n_states = 0, n_classes = 0, n_regex = 0;
    std::vector<uint16_t> classmap;      // [65536]  Automata._alphabet
    std::vector<int32_t> trans;          // [S*C]    Automata._transitions, -1 = dead
    std::vector<int32_t> accept_first;   // [S]      min(Automata._accept[s]) or -1
    std::vector<uint32_t> accept_off;    // [S+1]    CSR of Automata._accept
    std::vector<int32_t> accept_list;
};

DfaTables build_product(const std::vector<MinDfa>& dfas, size_t max_states = 4000000);

struct CompactDfa {                // GPU-oriented, same language + same first-accept map
    uint32_t n_states = 0, n_classes = 0;
    std::vector<uint16_t> classmap;      // [65536] -> compact class
    std::vector<int32_t> trans;          // [S'*C'], -1 = dead; state 0 = start
    std::vector<int32_t> accept_first;   // [S']
};

CompactDfa compact_tables(const DfaTables& t);

}  // namespace gorp
