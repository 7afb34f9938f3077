// One-pass extraction automaton (load-time, host only).
//
// The reference walks a line twice: PolyMatcher.match over the combined DFA (autom/PolyMatcher.java:123-133), then
// Pattern_e.matcher(line).matches() for the first accepting extraction e (jdkre/JDKRegexpCookedExtraction.java:36-39).
// For small definitions both walks are folded into ONE deterministic automaton so the GPU touches every unit once:
//
//   state   = (combined-DFA state q, capture-automaton state t_e for every extraction e that can still become the
//             FIRST accepting index from q)                                   -- exact product, no language assumption
//   symbol  = joint class of the unit under both class maps (+ the PAIR_HI look-ahead symbol of the capture side)
//   command = the register commands of all component transitions; the run time does not execute them but records
//             "the last position at which command list L fired" in one slot per distinct list ("op slot")
//   outcome = at end of line: MISS | MATCH(e, which op slots give each group boundary) | CAPTURE_FAIL(e)
//
// A capture register's final value is the position of the last transition that wrote it, i.e. the maximum over the
// op slots whose list writes it. That holds only for "register := position" commands, so the automaton is refused
// (FusedAutomaton::available == false, the engine then keeps the two-pass kernels) when a reachable capture
// transition carries a register copy, when a boundary has more than 4 writers, or when it outgrows the limits.
#pragma once
#include "model.hpp"

namespace gorp {

struct FusedAutomaton {
    bool available = false;
    std::string why_not;              // reason when !available

    // joint alphabet
    uint32_t n_jcls = 0;
    std::vector<uint16_t> jcls;       // [65536] unit -> joint class
    std::vector<uint16_t> pair_of;    // [J] joint class to use when the unit is a high surrogate FOLLOWED by a low one

    // automaton (minimised; state 0 = start)
    uint32_t n_states = 0;
    std::vector<uint32_t> trans;      // [S*J] low 16: next state (0xFFFF = dead => MISS), high 16: op slot (0 = none)
    uint32_t n_op_slots = 0;          // op slots are numbered 1..n_op_slots

    // end of line
    struct Outcome {
        int32_t ext_code;             // -1 MISS, e >= 0 MATCH, -2-e CAPTURE_FAIL (same coding as gorp_result.ext_id)
        uint32_t res_off;             // MATCH: first of 2*groups(e) entries in `res`
    };
    std::vector<uint32_t> outcome_of; // [S]
    std::vector<Outcome> outcomes;    // outcome 0 = MISS
    // group boundary k of a MATCH outcome = max over up to 4 op slots packed one per byte (0 terminates;
    // kLenSlot = the line length; no writer at all => -1, the group did not participate)
    std::vector<uint32_t> res;
    static constexpr uint32_t kLenSlot = 0xFF;
};

FusedAutomaton build_fused(const DeviceModel& m, size_t max_states = 2048, size_t max_op_slots = 120);

// Last load-time step: a definition that does not get the one-pass automaton takes the index -> DFA -> bucket -> capture
// path, whose per-extraction tables live in shared memory — their automata are minimised. (With the one-pass automaton
// the per-extraction automata only serve the List<String> form and stay as built.)
void finalize_device_model(DeviceModel& m, const FusedAutomaton& fused);

}  // namespace gorp
