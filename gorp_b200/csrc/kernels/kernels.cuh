// Device-side data layout and kernel entry points of libgorpcuda (sm_100a).
//
// Reference functions each kernel stands in for (gorp-core/src/main/java/com/salesforce/gorp/):
//   K1 newline index    — no counterpart (callers of Gorp.extract pre-split lines, Gorp.java:145)
//   K2 dfa_scan         — PolyMatcher.match + Automata.step/accept (autom/PolyMatcher.java:123-133,
//                         autom/Automata.java:133-139) and the first-index dispatch of Gorp.java:166-167
//   K4 tdfa_capture     — JDKRegexpCookedExtraction.match/_constructMatch (jdkre/JDKRegexpCookedExtraction.java:36-59)
//   K3 histogram, K5 span offsets (exclusive scan) — result assembly of model/CookedExtraction.java:54-57
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gorp {

constexpr int kMaxTdfaRegs = 32;   // run-time register file per line (tag registers of the capture automaton)

struct DfaDev {                    // combined multi-regex DFA, compacted (host/automata.hpp: CompactDfa)
    const uint16_t* cls;           // [65536] unit -> class
    const void* trans;             // [(S+1)*(C+1)] premultiplied next-row offset (next*(C+1)); u16 entries, or u32
                                   // when `wide`. Row S = dead (absorbing); column C = identity (self-loop).
    const int32_t* accept_first;   // [S+1], entry S = -1
    uint32_t n_states, n_classes;  // S, C (without the dead row / identity column)
    uint32_t wide;
};

struct ExtDev {                    // per-extraction capture automaton (host/capture.hpp: Tdfa)
    uint32_t n_states;             // S; row S of the table = dead (absorbing, not accepting)
    uint32_t trans_off;            // into tdfa_trans: [(S+1) * (n_classes+1)], last column = identity
    uint32_t opoff_off;            // into tdfa_op_off
    uint32_t ops_off;              // into tdfa_ops
    uint32_t fin_off;              // into tdfa_fin
    uint32_t acc_off;              // into tdfa_accepting
    uint32_t n_slots;              // 2 * groups
};

struct CapDev {
    const uint16_t* cls;           // [65536] unit -> symbol class of the capture automata
    uint32_t n_classes, pair_hi_class;
    const ExtDev* ext;             // [E]
    const uint32_t* tdfa_trans;    // next(16) | oplist(16)
    const uint32_t* tdfa_op_off;
    const uint16_t* tdfa_ops;      // (dst << 8) | src, src 0xFF = position
    const uint8_t* tdfa_fin;
    const uint8_t* tdfa_accepting;
    uint32_t n_ext;
    uint32_t match_only;
};

struct Launch {
    cudaStream_t stream;
    int sm_count;
};

// K1: '\n' index over UTF-16 text. tile_counts/tile_base sized ceil(n_units / kNlTile).
constexpr int kNlTile = 8192;
void k1_count_newlines(const Launch&, const uint16_t* text, int64_t n_units, uint32_t* tile_counts);
void k1_scatter_newlines(const Launch&, const uint16_t* text, int64_t n_units, const int64_t* tile_base,
                         int64_t* line_off /* entries 1.. */);
void k1_finish(const Launch&, const uint16_t* text, int64_t n_units, const int64_t* total_newlines, int64_t* line_off,
               int64_t* n_lines_out);

// exclusive scan: out[i] = sum_{j<i} in[j], out[n] = total (int64). scratch >= ceil(n/4096)+1 int64.
void scan_u32_to_i64(const Launch&, const uint32_t* in, int64_t n, int64_t* out, int64_t* scratch);

// K2: combined DFA, one line per thread. Writes ext_id (>=0 | -1) and span_cnt = 2*groups(ext) (0 on miss).
void k2_dfa_scan(const Launch&, const DfaDev&, const uint16_t* text, const int64_t* line_off, int sep, int64_t n_lines,
                 const uint32_t* slots_per_ext, int32_t* ext_id, uint32_t* span_cnt);

// K4: capture automaton over the matched lines. Writes spans at span_off[i]; capture failure => ext_id = -2-e, spans -1.
void k4_tdfa_capture(const Launch&, const CapDev&, const uint16_t* text, const int64_t* line_off, int sep, int64_t n_lines,
                     const int64_t* span_off, int32_t* ext_id, int32_t* spans);

// K3: per-extraction histogram (E entries, then MISS, then capture failures).
void k3_histogram(const Launch&, const int32_t* ext_id, int64_t n_lines, uint32_t n_ext, unsigned long long* hist);

}  // namespace gorp
