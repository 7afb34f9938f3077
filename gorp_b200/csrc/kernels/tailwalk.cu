// K4c — tail walk (sm_100a): the capture half of the text form for extractions that have a tail automaton
// (host/tails.hpp): F_e = minimal DFA of regex_e x capture automaton of e. One walk over the line decides the outcome of
// Gorp.extract for a line whose combined-DFA walk stopped early with candidate e (reference Gorp.java:159-177: MISS when
// regex_e rejects, CAPTURE_FAIL when only java.util.regex rejects, else MATCH) and yields the group spans of
// JDKRegexpCookedExtraction.match/_constructMatch (jdkre/JDKRegexpCookedExtraction.java:36-59).
//
// Work decomposition = kernels/capwalk.cu (lines bucketed by candidate extraction, a CTA takes work items of <= 4096
// lines of ONE extraction, lanes claim lines from the CTA's cursor one line ahead), the per-unit step = kernels/
// chunkwalk.cu: the extraction's table lives in shared memory as directly ASCII-indexed rows of u16 entries,
//
//   per unit :  ent = LDS.U16[row(ent) + 2*unit]          one shared-memory lookup, the only dependent chain
//               STS.U16 slot(ent)[thread] = position + 1    "last position at which command list `slot` fired"
//   per line :  the '\n' column leads to the absorbing row of the line's OUTCOME; a dead transition to the MISS outcome
//
// no class lookup, no trap rows, no register copies. Positions are line-relative and held as u16 (+1, 0 = not written):
// lines of kTailMaxLen units or more are handed to a one-thread-per-line kernel with 32-bit slots (tail_long_kernel).
// Result rows are written by the thread that walked the line, batched every `flush_every` iterations so that the lanes
// of a warp that have a finished line write together; two banks of slots keep the finished line's positions meanwhile.
#include "device_common.cuh"

namespace gorp {

namespace {

using namespace dev;

__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u16(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(static_cast<unsigned short>(v)) : "memory");
}

// one step on an ASCII unit held in byte kByte (0 or 2) of w. `ra` = byte offset of the current row + rows_abs.
template <int kByte>
__device__ __forceinline__ void tw_step(uint32_t& ra, uint32_t w, uint32_t rows_abs, uint32_t row_bytes, uint32_t slot_abs, uint32_t pos1) {
    uint32_t b, a;
    asm("prmt.b32 %0, %1, 0, %2;" : "=r"(b) : "r"(w), "n"(kByte == 0 ? 0x4440 : 0x4442));
    asm("mad.lo.u32 %0, %1, 2, %2;" : "=r"(a) : "r"(b), "r"(ra));
    const uint32_t ent = lds_u16(a);
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(ra) : "r"(ent >> 6), "r"(row_bytes), "r"(rows_abs));
    uint32_t sa;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(sa) : "r"(ent & 63u), "n"(kTailWalkThreads * 2), "r"(slot_abs));
    sts_u16(sa, pos1);
}

// 16 units that hold a unit >= 0x80: unit by unit through the column map (global, L1/L2 resident). A high surrogate
// followed by a low surrogate takes the PAIR column (java.util.regex consumes the pair as one character).
__device__ __noinline__ uint32_t tw_slow16(const TailDev& T, uint32_t ra, const Units16& u, const uint16_t* __restrict__ text, int64_t q,
                                           int64_t n_units, uint32_t rows_abs, uint32_t row_bytes, uint32_t fin_ra, uint32_t slot_abs,
                                           uint32_t pos1) {
    const uint32_t w[8] = {u.a.x, u.a.y, u.a.z, u.a.w, u.b.x, u.b.y, u.b.z, u.b.w};
#pragma unroll 1
    for (int k = 0; k < 16; ++k) {
        if (ra >= fin_ra) break;
        const uint32_t cu = (k & 1) ? (w[k >> 1] >> 16) : (w[k >> 1] & 0xFFFFu);
        uint32_t col = cu;
        if (cu >= 0x80u) {
            col = __ldg(T.xcol + cu);
            if ((cu & 0xFC00u) == 0xD800u) {
                const uint32_t nx = k == 15 ? (q + 16 < n_units ? __ldg(text + q + 16) : 0x0Au)
                                            : ((k & 1) ? (w[(k + 1) >> 1] & 0xFFFFu) : (w[k >> 1] >> 16));
                if ((nx & 0xFC00u) == 0xDC00u) col = __ldg(T.pair_col + col);
            }
        }
        const uint32_t ent = lds_u16(ra + col * 2);
        ra = (ent >> 6) * row_bytes + rows_abs;
        sts_u16(slot_abs + (ent & 63u) * (kTailWalkThreads * 2), pos1 + k);
    }
    return ra;
}

struct TwPending {  // a finished line whose result row has not been written yet
    uint32_t line, outcome, len, bank_abs;
};

__global__ void __launch_bounds__(kTailWalkThreads, 2) tailwalk_kernel(TailWalkParams P) {
    extern __shared__ __align__(16) unsigned char s_mem[];  // [table][recipes][outcome codes][init list][slots: 2 banks]
    __shared__ uint32_t s_item, s_cursor, s_loaded;
    const TailDev& T = P.t;
    const uint32_t stride = T.span_stride;
    const uint32_t tab_bytes = T.max_table_bytes;
    uint32_t* s_res = reinterpret_cast<uint32_t*>(s_mem + tab_bytes);
    int32_t* s_oext = reinterpret_cast<int32_t*>(s_res + T.max_res);
    uint8_t* s_init = reinterpret_cast<uint8_t*>(s_oext + T.max_outcomes);
    unsigned char* s_slots = reinterpret_cast<unsigned char*>(s_init + 64);
    const uint32_t rows_abs = static_cast<uint32_t>(__cvta_generic_to_shared(s_mem));
    const uint32_t slot_abs0 = static_cast<uint32_t>(__cvta_generic_to_shared(s_slots)) + threadIdx.x * 2;
    const uint32_t slot_stride = kTailWalkThreads * 2, bank_bytes = T.max_slots * slot_stride;
    const uint32_t row_bytes = T.row_bytes;
    const uint32_t lane = threadIdx.x & 31u, lt_mask = (1u << lane) - 1u;
    const uint32_t n_items = *P.n_items;
    const uint32_t flush_mask = P.flush_every - 1u;
    if (threadIdx.x == 0) s_loaded = 0xFFFFFFFFu;

    for (;;) {
        __syncthreads();  // the previous item is finished (its table and s_item / s_cursor are free)
        if (threadIdx.x == 0) s_item = atomicAdd(P.item_ticket, 1u);
        __syncthreads();
        const uint32_t item = s_item;
        if (item >= n_items) break;
        const CapItem it = P.items[item];
        const TailExt x = T.ext[it.ext];
        if (!x.available) continue;  // the bucketed capture walk (kernels/capwalk.cu) takes the items of this extraction
        if (s_loaded != it.ext) {
            const uint32_t n16 = ((x.fin_base + x.n_outcomes) * row_bytes + 15u) / 16u;
            const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(T.image) + x.tab_off);
            uint4* dst = reinterpret_cast<uint4*>(s_mem);
            for (uint32_t i = threadIdx.x; i < n16; i += kTailWalkThreads) dst[i] = __ldg(src + i);
            for (uint32_t i = threadIdx.x; i < x.n_outcomes * stride; i += kTailWalkThreads) {
                const uint32_t rec = __ldg(T.res + x.res_off + i);
                // one writer: byte offset of its slot inside a bank; several: bit 31 | one slot id per byte; 0 = none; 0xFF = length
                s_res[i] = (rec == 0u || rec == 0xFFu) ? rec << 24 : (rec < 256u ? rec * slot_stride : 0x80000000u | rec);
            }
            for (uint32_t i = threadIdx.x; i < x.n_outcomes; i += kTailWalkThreads) s_oext[i] = __ldg(T.oext + x.oext_off + i);
            for (uint32_t i = threadIdx.x; i < x.n_init; i += kTailWalkThreads) s_init[i] = __ldg(T.init_slots + x.init_off + i);
        }
        if (threadIdx.x == 0) s_cursor = it.begin;
        __syncthreads();
        if (threadIdx.x == 0) s_loaded = it.ext;

        const int32_t cand = static_cast<int32_t>(it.ext);
        const uint32_t fin_ra = x.fin_base * row_bytes + rows_abs;
        const uint32_t skip_ra = x.n_states * row_bytes + rows_abs;  // SKIP_1; SKIP_k = skip_ra + (k - 1) * row_bytes

        // writes the result row of a finished line (the thread's own; rows are `stride` int32)
        auto flush = [&](const TwPending& pd) {
            const int32_t code = s_oext[pd.outcome];
            if (code != cand) {  // the candidate did not match after all: MISS (regex_e rejects) or CAPTURE_FAIL
                P.ext_id[pd.line] = code;
                atomicAdd(P.hist + cand, ~0ull);  // -1
                atomicAdd(P.hist + P.n_ext + (code == -1 ? 0u : 1u), 1ull);
            }
            int32_t* out = P.spans + static_cast<int64_t>(pd.line) * stride;
            const uint32_t* res = s_res + pd.outcome * stride;
            auto value = [&](uint32_t recipe) -> int32_t {
                if (code < 0 || recipe == 0u) return -1;
                if (recipe == 0xFF000000u) return static_cast<int32_t>(pd.len);
                if (!(recipe & 0x80000000u)) return static_cast<int32_t>(lds_u16(pd.bank_abs + recipe)) - 1;
                recipe &= 0x7FFFFFFFu;
                uint32_t best = lds_u16(pd.bank_abs + (recipe & 0xFFu) * slot_stride);
                for (recipe >>= 8; recipe; recipe >>= 8) best = max(best, lds_u16(pd.bank_abs + (recipe & 0xFFu) * slot_stride));
                return static_cast<int32_t>(best) - 1;
            };
            if ((stride & 3u) == 0) {
                for (uint32_t k = 0; k < stride; k += 4) {
                    const uint4 r4 = *reinterpret_cast<const uint4*>(res + k);
                    *reinterpret_cast<int4*>(out + k) = make_int4(value(r4.x), value(r4.y), value(r4.z), value(r4.w));
                }
            } else if ((stride & 1u) == 0) {
                for (uint32_t k = 0; k < stride; k += 2) {
                    const uint2 r2 = *reinterpret_cast<const uint2*>(res + k);
                    *reinterpret_cast<int2*>(out + k) = make_int2(value(r2.x), value(r2.y));
                }
            } else {
                for (uint32_t k = 0; k < stride; ++k) out[k] = value(res[k]);
            }
        };

        // per-lane state: the line being walked, and the NEXT line of the lane, claimed one line ahead so that its id
        // (perm), start and end (line_off) are loaded long before they are needed
        bool exhausted = false;  // warp-uniform: the item has no unclaimed line left
        bool active = false, pending = false;
        TwPending pd{0, 0, 0, 0};
        uint32_t line = 0, len = 0, it_count = 0;
        int64_t a = 0, q = 0;
        uint32_t ra = fin_ra, bank_abs = slot_abs0;
        uint32_t nstage = 0;  // 0 = no next line, 1 = id requested, 2 = start and end requested
        uint32_t nline = 0;
        int64_t na = 0, nb = 0;
        for (;;) {
            if (!active && nstage) {  // start the claimed line
                if (nstage == 1) {
                    na = __ldg(P.line_off + nline);
                    nb = __ldg(P.line_off + nline + 1);
                }
                nstage = 0;
                const int64_t len64 = nb - 1 - na;
                if (len64 >= static_cast<int64_t>(kTailMaxLen)) {  // 32-bit positions: tail_long_kernel
                    const uint32_t slot = atomicAdd(P.n_long, 1u);
                    if (slot < P.long_cap) P.long_lines[slot] = nline;
                } else {
                    line = nline;
                    a = na;
                    len = static_cast<uint32_t>(len64);
                    q = a & ~int64_t(15);
                    const uint32_t lo = static_cast<uint32_t>(a - q);
                    ra = lo ? skip_ra + (lo - 1) * row_bytes : rows_abs;
                    if (pending) bank_abs = pd.bank_abs == slot_abs0 ? slot_abs0 + bank_bytes : slot_abs0;
                    for (uint32_t i = 0; i < x.n_init; ++i) sts_u16(bank_abs + s_init[i] * slot_stride, 0u);
                    active = true;
                }
            } else if (nstage == 1) {
                na = __ldg(P.line_off + nline);
                nb = __ldg(P.line_off + nline + 1);
                nstage = 2;
            }
            if (!exhausted) {  // lanes without a next line claim the next entries of the item
                const uint32_t want = __ballot_sync(0xffffffffu, nstage == 0);
                if (want) {
                    uint32_t base = 0;
                    if (lane == 0) base = atomicAdd(&s_cursor, static_cast<uint32_t>(__popc(want)));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    const uint32_t idx = base + static_cast<uint32_t>(__popc(want & lt_mask));
                    if (nstage == 0 && idx < it.end) {
                        nline = __ldg(P.perm + idx);
                        nstage = 1;
                    }
                    exhausted = base + static_cast<uint32_t>(__popc(want)) >= it.end;
                }
            }
            if (!__any_sync(0xffffffffu, active || nstage != 0)) break;
            if (active) {
                const Units16 u = load_units16_l2keep(P.text, q, P.n_units);
                const uint32_t pos1 = static_cast<uint32_t>(q - a) + 1u;  // garbage while skipping: only ever stored to the dummy slot
                if (((u.a.x | u.a.y | u.a.z | u.a.w | u.b.x | u.b.y | u.b.z | u.b.w) & 0xFF80FF80u) == 0u) {
                    tw_step<0>(ra, u.a.x, rows_abs, row_bytes, bank_abs, pos1);
                    tw_step<2>(ra, u.a.x, rows_abs, row_bytes, bank_abs, pos1 + 1);
                    tw_step<0>(ra, u.a.y, rows_abs, row_bytes, bank_abs, pos1 + 2);
                    tw_step<2>(ra, u.a.y, rows_abs, row_bytes, bank_abs, pos1 + 3);
                    tw_step<0>(ra, u.a.z, rows_abs, row_bytes, bank_abs, pos1 + 4);
                    tw_step<2>(ra, u.a.z, rows_abs, row_bytes, bank_abs, pos1 + 5);
                    tw_step<0>(ra, u.a.w, rows_abs, row_bytes, bank_abs, pos1 + 6);
                    tw_step<2>(ra, u.a.w, rows_abs, row_bytes, bank_abs, pos1 + 7);
                    tw_step<0>(ra, u.b.x, rows_abs, row_bytes, bank_abs, pos1 + 8);
                    tw_step<2>(ra, u.b.x, rows_abs, row_bytes, bank_abs, pos1 + 9);
                    tw_step<0>(ra, u.b.y, rows_abs, row_bytes, bank_abs, pos1 + 10);
                    tw_step<2>(ra, u.b.y, rows_abs, row_bytes, bank_abs, pos1 + 11);
                    tw_step<0>(ra, u.b.z, rows_abs, row_bytes, bank_abs, pos1 + 12);
                    tw_step<2>(ra, u.b.z, rows_abs, row_bytes, bank_abs, pos1 + 13);
                    tw_step<0>(ra, u.b.w, rows_abs, row_bytes, bank_abs, pos1 + 14);
                    tw_step<2>(ra, u.b.w, rows_abs, row_bytes, bank_abs, pos1 + 15);
                } else {
                    ra = tw_slow16(T, ra, u, P.text, q, P.n_units, rows_abs, row_bytes, fin_ra, bank_abs, pos1);
                }
                q += 16;
                if (ra >= fin_ra) {  // the line ended inside these 16 units (its '\n', a dead transition, or the end of the text)
                    if (pending) flush(pd);  // rare: two line ends of one lane between two flush points
                    pd.line = line;
                    pd.outcome = (ra - fin_ra) / row_bytes;
                    pd.len = len;
                    pd.bank_abs = bank_abs;
                    pending = true;
                    active = false;
                }
            }
            if ((++it_count & flush_mask) == 0 && pending) {
                flush(pd);
                pending = false;
            }
        }
        if (pending) flush(pd);
    }
}

// Lines of kTailMaxLen units or more: one thread per line, table in global memory, 32-bit slots in local memory.
__global__ void __launch_bounds__(32) tail_long_kernel(TailWalkParams P) {
    const TailDev& T = P.t;
    const uint32_t n = min(*P.n_long, P.long_cap);
    const uint32_t stride = T.span_stride, width = T.width;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t line = P.long_lines[i];
        const int32_t cand = P.ext_id[line];
        if (cand < 0) continue;
        const TailExt x = T.ext[cand];
        const uint16_t* __restrict__ tab = reinterpret_cast<const uint16_t*>(reinterpret_cast<const unsigned char*>(T.image) + x.tab_off);
        uint32_t slots[64];
        for (uint32_t k = 0; k < 64; ++k) slots[k] = 0;
        const int64_t a = P.line_off[line], b = P.line_off[line + 1] - 1;
        uint32_t row = 0;
        for (int64_t p = a; row < x.fin_base; ++p) {
            const uint32_t cu = p < b && p < P.n_units ? P.text[p] : 0x0Au;  // position b holds the '\n' (or lies beyond the text)
            uint32_t col = cu;
            if (cu >= 0x80u) {
                col = T.xcol[cu];
                if ((cu & 0xFC00u) == 0xD800u && p + 1 < b && (P.text[p + 1] & 0xFC00u) == 0xDC00u) col = T.pair_col[col];
            }
            const uint32_t ent = tab[static_cast<size_t>(row) * width + col];
            row = ent >> 6;
            slots[ent & 63u] = static_cast<uint32_t>(p - a) + 1u;
        }
        const uint32_t o = row - x.fin_base;
        const int32_t code = T.oext[x.oext_off + o];
        if (code != cand) {
            P.ext_id[line] = code;
            atomicAdd(P.hist + cand, ~0ull);
            atomicAdd(P.hist + P.n_ext + (code == -1 ? 0u : 1u), 1ull);
        }
        int32_t* out = P.spans + static_cast<int64_t>(line) * stride;
        for (uint32_t k = 0; k < stride; ++k) {
            uint32_t rec = T.res[x.res_off + o * stride + k];
            int32_t val = -1;
            if (code >= 0 && rec == 0xFFu) {
                val = static_cast<int32_t>(b - a);
            } else if (code >= 0 && rec) {
                uint32_t best = 0;
                for (; rec; rec >>= 8) best = max(best, slots[rec & 0xFFu]);
                val = static_cast<int32_t>(best) - 1;
            }
            out[k] = val;
        }
    }
}

}  // namespace

size_t tailwalk_smem_bytes(const TailDev& t) {
    return static_cast<size_t>(t.max_table_bytes) + static_cast<size_t>(t.max_res) * 4 + static_cast<size_t>(t.max_outcomes) * 4 + 64 +
           2 * static_cast<size_t>(t.max_slots) * kTailWalkThreads * 2;
}

void k4c_tailwalk(const Launch& L, const TailWalkParams& P) {
    const size_t smem = tailwalk_smem_bytes(P.t);
    allow_max_dynamic_smem(tailwalk_kernel);
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tailwalk_kernel, kTailWalkThreads, smem);
    if (per_sm < 1) per_sm = 1;
    tailwalk_kernel<<<L.sm_count * per_sm, kTailWalkThreads, smem, L.stream>>>(P);
    tail_long_kernel<<<L.sm_count, 32, 0, L.stream>>>(P);
}

}  // namespace gorp
