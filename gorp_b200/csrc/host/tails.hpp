// Per-extraction "tail" automata and the early-exit cut of the combined DFA (load-time, host only).
//
// The reference decides a line in two full walks: PolyMatcher.match over the combined DFA
// (autom/PolyMatcher.java:123-133), then Pattern_e.matcher(line).matches() for the first accepting extraction e
// (jdkre/JDKRegexpCookedExtraction.java:36-39). For big definitions the combined DFA is large (config #4: 11 115 states)
// but almost all of its states can only ever lead to ONE extraction: let Alive(q) be the set of extractions that can
// still become the first accepting index from state q. Once the walk reaches a state with Alive(q) = {e}
//   * no extraction below e accepts any continuation (it would be a possible first index), so
//   * the line's outcome is MISS unless regex_e (the DFA dialect string of e) accepts the whole line, and then it is
//     MATCH(e, spans) or CAPTURE_FAIL(e) depending on the java.util.regex side — exactly Gorp.extract (Gorp.java:159-177).
// So the combined-DFA walk may stop there ("cut"), and the capture pass walks the line once over
//   F_e = minimal DFA of regex_e  x  capture automaton of e        (exact product, minimised)
// which yields the three outcomes and the spans in one lookup per unit. The combined DFA shrinks to its states with
// |Alive| >= 2 (config #4: 31 states) and fits shared memory; F_e has the size of the capture automaton.
//
// regex_e's DFA is recovered from the tables the blob carries (Automata._transitions/_accept, reference
// autom/Automata.java:23-26): the language of e = the product automaton with acceptance "e in _accept[state]", trimmed
// and minimised — the automaton-dialect string itself is not retained by the reference (see java/.../DfaExport.java).
//
// Format of a tail = the one-pass automaton of host/fused.hpp (op slots = "last position at which command list L fired",
// outcomes with per-boundary writer slots), with transitions laid out by COLUMN:
//   columns [0,128)  ASCII units (column 0x0A is never used by a '\n'-terminated line: the terminator)
//   columns >= 128   shared by all tails: one per (DFA class, capture class) pair that a unit >= 0x80 can take, the
//                    PAIR_HI variants (high surrogate followed by a low surrogate), and one for a '\n' that is line
//                    CONTENT (List<String> form)
#pragma once
#include "fused.hpp"

namespace gorp {

struct TailAutomaton {
    bool available = false;
    std::string why_not;
    uint32_t n_states = 0;
    std::vector<uint32_t> trans;       // [S * width] low 16: next state (0xFFFF = dead => MISS), high 16: op slot (0 = none)
    uint32_t n_op_slots = 0;           // op slots are numbered 1..n_op_slots
    uint32_t n_boundaries = 0;         // 2 * groups of the extraction
    std::vector<uint32_t> outcome_of;  // [S]
    std::vector<FusedAutomaton::Outcome> outcomes;  // outcome 0 = MISS; ext_code = -1 | e | -2-e
    std::vector<uint32_t> res;         // per MATCH outcome: 2*groups entries, up to 4 writer slots packed one per byte
                                       // (0 terminates, FusedAutomaton::kLenSlot = the line length)
    std::vector<uint32_t> init_slots;  // slots that are read through a several-writer maximum: reset per line
};

struct TailSet {
    bool any = false;                  // at least one extraction has a tail
    uint32_t width = 0;                // columns per row (multiple of 4)
    std::vector<uint16_t> xcol;        // [65536] unit -> column (units < 0x80 map to themselves)
    std::vector<uint16_t> pair_col;    // [width] column of a high surrogate that is followed by a low surrogate
    uint32_t nl_data_col = 0;          // column of a '\n' inside a line (List<String> form)
    std::vector<TailAutomaton> tails;  // [E]
    std::vector<int32_t> cut_of_state; // [compact DFA states] e when the walk may stop here and hand the line to tail e, else -1
    uint32_t n_cut_states = 0, n_head_states = 0;
};

// `max_states` / `max_op_slots`: per tail; a tail beyond the limits is not available (its extraction keeps the two-walk path
// and none of its states is cut).
TailSet build_tails(const CompiledDefinition& def, const DeviceModel& m, size_t max_states = 1000, size_t max_op_slots = 62);

// The one-pass automaton of host/fused.hpp laid out as ONE tail (the "fused walk" of small definitions: every line walks it,
// behind the newline index): TailSet with a single automaton whose outcomes carry the codes of all extractions; recipes are
// padded to `span_stride` entries per outcome. Columns >= 128 are the joint classes of units >= 0x80, their PAIR_HI variants
// and the '\n'-as-content column.
TailSet build_fused_tailset(const FusedAutomaton& A, const DeviceModel& m, uint32_t span_stride);

}  // namespace gorp
