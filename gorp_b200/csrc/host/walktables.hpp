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

namespace gorp {

struct DfaWalkTable {
    bool available = false;
    uint32_t n_rows = 0, K = 0, n_states = 0, n_classes = 0, fin_base = 0;
    std::vector<uint16_t> rows;    // [n_rows * K] next row, padded to 16 bytes
    std::vector<uint16_t> cls128;  // [128] ASCII unit -> 2 * column ('\n' -> the '\n' column)
    std::vector<uint16_t> xcls;    // [65536] unit -> column
};
DfaWalkTable build_dfawalk_table(const DeviceModel& m);

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

}  // namespace gorp
