// Tables of the big-definition text path (load-time, host only; uploaded by engine.cu, walked by kernels/dfawalk.cu
// and kernels/capwalk.cu, interpreted on the CPU by tests/host/hosttest.cpp).
//
// DfaWalkTable — the combined DFA (reference autom/Automata.java:133-139 step/accept) as class-indexed u16 rows:
//   rows [0,S) states | S = DEADSCAN (dead, keeps scanning to the line's '\n') | S+1..S+15 = SKIP_1..15 (swallow the units
//   that precede the line in its first 32-byte block) | fin_base = S+16: FIN(-1), then FIN(0..E-1), absorbing.
//   Columns: the classes ordered by how often log text hits them, then '\n' (column n_classes), then padding.
// CapImage — one table per extraction for the capture automaton (host/capture.hpp: Tdfa), K u32 entries per row:
//   rows [0,S) states | S..S+14 = SKIP_1..15 | S+15 = DEAD | S+16 = SLOW (trap: the block is replayed through the general
//   tables) | S+17+s = FRZ(s), reached at the line's '\n' from state s, absorbing.
//   Entry = (byte offset of the next row inside the extraction's table) << 6 | register slot set to the current position
//   (slot n_regs = the per-thread dummy).
#pragma once
#include "model.hpp"
#include "tails.hpp"

namespace gorp {

struct DfaWalkTable {
    bool available = false;
    uint32_t n_rows = 0, K = 0, n_states = 0, n_classes = 0, fin_base = 0;
    std::vector<uint16_t> rows;    // [n_rows * K] next row, padded to 16 bytes
    std::vector<uint16_t> cls128;  // [128] ASCII unit -> 2 * column ('\n' -> the '\n' column)
    std::vector<uint16_t> xcls;    // [65536] unit -> column
};
DfaWalkTable build_dfawalk_table(const DeviceModel& m);
// The same table with the early-exit cut of host/tails.hpp, for walkers that know where a line ends (K2b): only the
// states reachable without passing a cut state are kept (renumbered), a transition into a cut state leads to
// FIN(candidate extraction) at once, a dead transition to FIN(-1) at once.
DfaWalkTable build_dfawalk_table_cut(const DeviceModel& m, const std::vector<int32_t>& cut_of_state);

struct CapImageExt {   // byte offsets (mirrors kernels.cuh: CapImgExt)
    uint32_t tab_off, row_bytes, n_states, dead_off, slow_off, frz_off;
};
struct CapImage {
    bool available = false;
    uint32_t K = 0, n_regs = 0;
    std::vector<uint32_t> image;   // all tables, each 16-byte aligned
    std::vector<uint32_t> cls128;  // [128] ASCII unit -> 4 * class ('\n' -> the last real column)
    std::vector<CapImageExt> ext;  // [E]
};
CapImage build_cap_image(const DeviceModel& m, size_t max_extractions);

// TailImage — the tail automata of host/tails.hpp as the capture pass walks them (kernels/tailwalk.cu), one table per
// extraction, `width` u16 entries per row:
//   rows [0,S) states | S..S+14 = SKIP_1..15 (swallow the units that precede the line in its first 32-byte block) |
//   fin_base = S+15: outcome rows (absorbing; outcome 0 = MISS)
//   columns: [0,128) ASCII units, column 0x0A = the line terminator (leads to the state's outcome row); >= 128 see tails.hpp
//   entry = (next row << 6) | op slot (0 = dummy; the walk stores position + 1 into the slot, 0 = "not written")
struct TailImageExt {  // mirrors kernels.cuh: TailExt
    uint32_t tab_off;      // byte offset of the table inside the image (16-byte aligned)
    uint32_t n_states, fin_base, n_outcomes;
    uint32_t res_off;      // first recipe of outcome 0 in `res` (span_stride recipes per outcome)
    uint32_t oext_off;     // first entry of `oext`
    uint32_t init_off, n_init;  // slots reset to "not written" when a line starts (read through a several-writer maximum)
    uint32_t n_slots;      // op slots + the dummy
    uint32_t available;
};
struct TailImage {
    bool available = false;
    uint32_t width = 0, row_bytes = 0, span_stride = 0;
    uint32_t max_table_bytes = 0, max_slots = 0;
    std::vector<uint16_t> image;
    std::vector<TailImageExt> ext;   // [E]
    std::vector<uint32_t> res;       // recipe: 0 = no writer (-1) | 0xFF = the line length | up to 4 slot ids, one per byte
    std::vector<int32_t> oext;       // ext code per outcome: -1 | e | -2-e
    std::vector<uint8_t> init_slots;
};
TailImage build_tail_image(const TailSet& T, uint32_t span_stride);

}  // namespace gorp
