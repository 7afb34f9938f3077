// Capture half of the definition compiler (load-time, host only).
//
// Input: the JDK-dialect regex string of one extraction (what reference jdkre/JDKRegexpExtractionCooker.java:23
// hands to Pattern.compile). Output:
//   1. CaptureProgram — a priority-ordered Pike-VM program (CHAR/SET/SPLIT/JMP/SAVE/MATCH) whose thread order
//      is java.util.regex's backtracking preference order; acceptance only at end of line (Matcher.matches()).
//   2. Tdfa — that program determinised with tag registers (lazy, one-symbol-lookahead TDFA): the state is the
//      ordered list of live Pike threads, transitions carry register copy/set commands. One table lookup per
//      UTF-16 unit on the GPU instead of a thread-list simulation; spans are identical by construction.
//
// Code points: java.util.regex CharProperty nodes (negated classes, ranges, '.', \D \S \W) consume a whole
// surrogate pair. The automata run over UTF-16 units with ONE extra input symbol, PAIR_HI = "high surrogate
// that is followed by a low surrogate"; code-point sets accept PAIR_HI (+ the following unit) iff they contain
// the supplementary planes, unit-level nodes never accept it.
#pragma once
#include "common.hpp"

namespace gorp {

using Ranges = std::vector<std::pair<uint32_t, uint32_t>>;

enum : uint8_t { OP_CHAR = 0, OP_SET = 1, OP_ANY = 2, OP_PAIRHI = 3, OP_SPLIT = 4, OP_JMP = 5, OP_SAVE = 6, OP_MATCH = 7 };

struct CaptureProgram {
    struct Inst { uint8_t op; int32_t a, b; };
    std::vector<Inst> insts;
    std::vector<Ranges> sets;  // BMP unit ranges; never applied to a PAIR_HI symbol
    int n_groups = 0;          // SAVE slots are 2*(g-1), 2*(g-1)+1 for group g >= 1
};

// Throws DefinitionParseError for what Pattern.compile rejects, UnsupportedError for constructs whose
// java.util.regex meaning the GPU path does not reproduce (see DESIGN.md "supported subset").
CaptureProgram compile_jdk_regex(const ustring& regex);

struct SymbolClasses {
    std::vector<uint16_t> classmap;  // [65536] unit -> class
    uint32_t n_classes = 0;          // including pair_hi_class
    uint32_t pair_hi_class = 0;
};
SymbolClasses build_symbol_classes(const std::vector<CaptureProgram>& progs);

struct Tdfa {
    uint32_t n_states = 0, n_classes = 0, n_regs = 0, n_slots = 0;
    std::vector<uint32_t> trans;     // [S*C]  low 16 = next state (0xFFFF dead), high 16 = op-list id (0 = none)
    std::vector<uint32_t> op_off;    // CSR: op-list id -> [op_off[id], op_off[id+1])
    std::vector<uint16_t> ops;       // (dst << 8) | src ; src 0xFF = current position
    std::vector<uint8_t> accepting;  // [S]
    std::vector<uint8_t> fin;        // [S*n_slots] register id | 0xFE = end position | 0xFF = unset (group null)
};

// Throws UnsupportedError when the determinisation exceeds the limits.
Tdfa build_tdfa(const CaptureProgram& p, const SymbolClasses& sc, size_t max_states = 60000, size_t max_regs = 250);

// A capture automaton that accepts nothing: stands in for an extraction whose determinisation exceeded the limits. Every
// line it is asked about ends as CAPTURE_FAIL on the table-driven paths; the simulating Pike-VM pass (PikeTables,
// kernels/pike.cu) then decides those lines for real.
Tdfa placeholder_tdfa(const CaptureProgram& p, const SymbolClasses& sc);

// The Pike program flattened for simulation (no determinisation): per instruction its epsilon closure in priority order
// (consuming target or MATCH, SAVE slots passed on the way) and, per symbol class, whether the instruction accepts it.
struct PikeTables {
    uint32_t n_insts = 0, n_slots = 0, n_classes = 0;
    std::vector<uint32_t> clo_off;    // [n_insts + 1]
    std::vector<int32_t> clo_target;  // consuming instruction index, -1 = MATCH
    std::vector<uint64_t> clo_mask;   // SAVE slots set to the current position on the way
    std::vector<uint8_t> accepts;     // [n_insts * n_classes]
};
PikeTables build_pike_tables(const CaptureProgram& p, const SymbolClasses& sc);

// Merges the states of `t` that no input suffix can tell apart (same acceptance, final recipe, command lists and equivalent
// successors). Used for the per-extraction tables of the bucketed capture walk; the one-pass product automaton
// (host/fused.hpp) is minimised as a whole and is built from the unminimised automata.
void minimise_tdfa(Tdfa& t);

}  // namespace gorp
