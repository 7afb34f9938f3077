// K0 — the fused, persistent extraction kernel (sm_100a): ONE pass over the text in HBM.
//
// Each CTA repeatedly takes the next 16 K-unit tile of text (in-order ticket), pulls it into shared memory with a
// 1-D TMA bulk copy (cp.async.bulk + mbarrier complete_tx), and then does everything the reference does per line
// for the lines that START after a '\n' of that tile (tile 0 also owns the line at offset 0):
//   A  newline discovery       (128-bit shared loads, __vcmpeq2, block scan)        — K1 of the unfused pipeline
//   B  combined DFA            (directly ASCII-indexed rows, 1 shared lookup/unit)  — PolyMatcher.match
//   C  span counts + block scan
//   D  decoupled look-back     (line and span prefixes across tiles, single pass)   — replaces the global scans
//   E  capture automaton       (class-indexed TDFA rows, tag registers in smem)     — Matcher.matches()+group(i)
//   F  coalesced result rows   (ext_id, line_off, span_off) + histogram
// Lines that run past the staged window (2 K units of overhang) continue from global memory; chunks with units
// >= 0x80 and transitions with more than one register command are replayed by the slow helpers in fast.cu.
// Tiles with more than kMaxLines line starts (average line < 16 units) raise FLAG_FALLBACK and the host reruns the
// batch through the unfused kernels; output capacity is checked against cap_lines/cap_spans (FLAG_OVERFLOW).
#include "device_common.cuh"

namespace gorp {

namespace {

constexpr int kT = kFusedThreads;
constexpr int kTile = kFusedTile;          // units per tile
constexpr int kOver = 2048;                // staged overhang (units)
constexpr int kBuf = kTile + kOver;        // staged units per tile
constexpr int kMaxLines = kFusedMaxLines;  // line starts per tile handled in shared memory

constexpr unsigned long long kFlagAgg = 1ull << 62, kFlagPre = 2ull << 62, kValMask = (1ull << 62) - 1;

using namespace dev;

// ---- mbarrier + TMA bulk copy (global -> shared::cta), raw PTX
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}

// chunk of 8 units at unit position `pos` of the text, read from global memory with units >= n_units replaced by '\n'
__device__ __noinline__ uint4 load_chunk_global(const uint16_t* __restrict__ text, int64_t pos, int64_t n_units) {
    if (pos + 8 <= n_units) return __ldg(reinterpret_cast<const uint4*>(text + pos));
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int64_t p0 = pos + 2 * j, p1 = p0 + 1;
        const uint32_t lo = p0 < n_units ? __ldg(text + p0) : 0x0Au;
        const uint32_t hi = p1 < n_units ? __ldg(text + p1) : 0x0Au;
        w[j] = lo | (hi << 16);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// block-wide exclusive scan of one value per thread (kT threads); returns exclusive prefix, total in *total
__device__ __forceinline__ uint32_t block_scan(uint32_t v, uint32_t* s_warp, uint32_t* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    __syncthreads();  // s_warp reuse
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kT / 32; ++w) {
        const uint32_t x = s_warp[w];
        if (w < warp) base += x;
        tot += x;
    }
    *total = tot;
    return base + incl - v;
}

__device__ __forceinline__ unsigned long long ld_acquire(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(kT, 2) fused_extract_kernel(FusedParams P) {
    extern __shared__ __align__(128) unsigned char smem[];
    // ---- carve shared memory
    uint32_t* s_rows = reinterpret_cast<uint32_t*>(smem);                                   // DFA rows
    uint32_t* s_img = s_rows + P.dfa.n_rows * 128;                                          // capture image
    uint32_t* s_regs = s_img + P.cap_fast.image_words;                                      // (n_regs + 2) * kT
    uint16_t* s_text = reinterpret_cast<uint16_t*>(s_regs + (P.cap_fast.n_regs + 2) * kT);  // kBuf + 8 units
    uint16_t* s_start = s_text + kBuf + 8;                                                  // kMaxLines
    int32_t* s_ext = reinterpret_cast<int32_t*>(s_start + kMaxLines);                       // kMaxLines
    uint32_t* s_spoff = reinterpret_cast<uint32_t*>(s_ext + kMaxLines);                     // kMaxLines
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ uint32_t s_warp[kT / 32];
    __shared__ uint32_t s_hist[kFusedHistBins];
    __shared__ long long s_tile, s_line_base, s_span_base;
    __shared__ int s_skip_writes;

    const uint32_t rows_abs = static_cast<uint32_t>(__cvta_generic_to_shared(s_rows));
    const uint32_t cls_abs = static_cast<uint32_t>(__cvta_generic_to_shared(s_img));
    const uint32_t text_abs = static_cast<uint32_t>(__cvta_generic_to_shared(s_text));
    const uint32_t bar = static_cast<uint32_t>(__cvta_generic_to_shared(&s_bar));
    const uint32_t reg_stride = kT * 4;
    const uint32_t reg_abs = static_cast<uint32_t>(__cvta_generic_to_shared(s_regs)) + threadIdx.x * 4;
    const uint32_t len_off = (P.cap_fast.n_regs + 1) * reg_stride;

    for (uint32_t i = threadIdx.x; i < P.dfa.n_rows * 128u; i += kT) s_rows[i] = rows_abs + (__ldg(P.dfa.rows + i) << 9);
    for (uint32_t i = threadIdx.x; i < P.cap_fast.image_words; i += kT) s_img[i] = __ldg(P.cap_fast.image + i);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const uint32_t fin_abs = rows_abs + (P.dfa.fin_base << 9);
    const uint32_t skip_abs = rows_abs + (P.dfa.skip_base << 9);
    const uint32_t n_bins = P.n_ext + 2;
    const bool smem_hist = n_bins <= kFusedHistBins;
    uint32_t parity = 0;
    __syncthreads();

    for (;;) {
        if (threadIdx.x == 0) s_tile = atomicAdd(P.ticket, 1u);
        __syncthreads();
        const int64_t tile = s_tile;
        if (tile >= P.n_tiles) break;
        const int64_t t0 = tile * kTile;
        // ---- stage [t0, t0 + kBuf) ∩ text into shared memory; pad the remainder with '\n'
        const int64_t avail = P.n_units - t0 < kBuf ? P.n_units - t0 : kBuf;  // > 0
        const uint32_t bulk_units = static_cast<uint32_t>(avail) & ~7u;
        if (threadIdx.x == 0) {
            if (bulk_units) {
                // the previous tile's generic-proxy reads of the buffer are ordered before this async-proxy write
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(bar, bulk_units * 2);
                tma_bulk_g2s(text_abs, P.text + t0, bulk_units * 2, bar);
            }
        }
        for (uint32_t i = bulk_units + threadIdx.x; i < static_cast<uint32_t>(kBuf + 8); i += kT)
            s_text[i] = i < avail ? __ldg(P.text + t0 + i) : static_cast<uint16_t>(0x0A);
        if (smem_hist)
            for (uint32_t i = threadIdx.x; i < n_bins; i += kT) s_hist[i] = 0;
        if (bulk_units) {
            mbar_wait(bar, parity);
            parity ^= 1;
        }
        __syncthreads();

        // ---- A: line starts owned by this tile = (position of '\n' in [t0, t0+kTile)) + 1, if < n_units
        constexpr int kPer = kTile / kT;  // consecutive units per thread (multiple of 8)
        uint32_t nl_mask = 0;             // kPer <= 32
        {
            const uint32_t u0 = threadIdx.x * kPer;
#pragma unroll
            for (int j = 0; j < kPer / 8; ++j) {
                const uint4 v = lds128(text_abs + (u0 + j * 8) * 2);
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t m = __vcmpeq2(w[k], 0x000A000Au);
                    nl_mask |= ((m & 1u) | ((m >> 15) & 2u)) << (j * 8 + k * 2);
                }
            }
            // a '\n' at position p starts a line only if p + 1 < n_units (padding beyond the text is '\n' too)
            const int64_t room = P.n_units - 1 - (t0 + u0);  // number of positions p in this thread's span with p+1 < n_units
            if (room < kPer) nl_mask = room <= 0 ? 0u : (nl_mask & ((1u << room) - 1u));
        }
        const uint32_t extra = tile == 0 ? 1u : 0u;  // the line at offset 0
        uint32_t n_t;
        uint32_t my = block_scan(__popc(nl_mask), s_warp, &n_t) + extra;
        n_t += extra;
        const bool too_dense = n_t > static_cast<uint32_t>(kMaxLines);
        if (!too_dense) {
            if (extra && threadIdx.x == 0) s_start[0] = 0;
            const uint32_t u0 = threadIdx.x * kPer;
            while (nl_mask) {
                const int k = __ffs(nl_mask) - 1;
                nl_mask &= nl_mask - 1;
                s_start[my++] = static_cast<uint16_t>(u0 + k + 1);  // 1 .. kTile
            }
        }
        __syncthreads();

        // ---- B: combined DFA per owned line
        if (!too_dense) {
            for (uint32_t i = threadIdx.x; i < n_t; i += kT) {
                const uint32_t rel = s_start[i];
                uint32_t q = rel & ~7u;
                const uint32_t lo = rel & 7u;
                uint32_t st = lo ? skip_abs + ((lo - 1) << 9) : rows_abs;
                do {
                    const uint4 v = q < static_cast<uint32_t>(kBuf) ? lds128(text_abs + q * 2) : load_chunk_global(P.text, t0 + q, P.n_units);
                    q += 8;
                    if (((v.x | v.y | v.z | v.w) & 0xFF80FF80u) == 0u) {
                        st = dfa_step<0>(st, v.x);
                        st = dfa_step<2>(st, v.x);
                        st = dfa_step<0>(st, v.y);
                        st = dfa_step<2>(st, v.y);
                        st = dfa_step<0>(st, v.z);
                        st = dfa_step<2>(st, v.z);
                        st = dfa_step<0>(st, v.w);
                        st = dfa_step<2>(st, v.w);
                    } else {
                        st = dfa_slow_chunk(P.dfa, rows_abs, st, v);
                    }
                } while (st < fin_abs);
                s_ext[i] = static_cast<int32_t>((st - fin_abs) >> 9) - 1;
            }
        }
        __syncthreads();

        // ---- C: span counts (2*groups per matched line) and their exclusive scan within the tile
        uint32_t span_t = 0;
        if (!too_dense) {
            constexpr int kLp = kMaxLines / kT;  // lines per thread in the scan
            uint32_t cnt[kLp];
            uint32_t sum = 0;
#pragma unroll
            for (int j = 0; j < kLp; ++j) {
                const uint32_t i = threadIdx.x * kLp + j;
                const int32_t e = i < n_t ? s_ext[i] : -1;
                cnt[j] = e >= 0 ? __ldg(P.slots_per_ext + e) : 0u;
                sum += cnt[j];
            }
            uint32_t run = block_scan(sum, s_warp, &span_t);
#pragma unroll
            for (int j = 0; j < kLp; ++j) {
                const uint32_t i = threadIdx.x * kLp + j;
                if (i < n_t) s_spoff[i] = run;
                run += cnt[j];
            }
        }

        // ---- D: decoupled look-back over tiles for (lines, spans)
        if (threadIdx.x == 0) {
            const unsigned long long my_lines = too_dense ? 0ull : n_t, my_spans = span_t;
            unsigned long long pre_l = 0, pre_s = 0;
            if (tile > 0) {
                st_release(P.tile_lines + tile, kFlagAgg | my_lines);
                st_release(P.tile_spans + tile, kFlagAgg | my_spans);
                for (int64_t j = tile - 1;; --j) {
                    unsigned long long a, b;
                    do {
                        a = ld_acquire(P.tile_lines + j);
                        b = ld_acquire(P.tile_spans + j);
                    } while ((a >> 62) == 0 || (b >> 62) == 0 || (a >> 62) != (b >> 62));
                    pre_l += a & kValMask;
                    pre_s += b & kValMask;
                    if ((a >> 62) == 2) break;
                }
            }
            st_release(P.tile_lines + tile, kFlagPre | (pre_l + my_lines));
            st_release(P.tile_spans + tile, kFlagPre | (pre_s + my_spans));
            s_line_base = static_cast<long long>(pre_l);
            s_span_base = static_cast<long long>(pre_s);
            int skip = too_dense ? 1 : 0;
            if (too_dense) atomicOr(reinterpret_cast<unsigned long long*>(P.totals + 2), 2ull);
            if (static_cast<long long>(pre_l + my_lines) > P.cap_lines || static_cast<long long>(pre_s + my_spans) > P.cap_spans) {
                atomicOr(reinterpret_cast<unsigned long long*>(P.totals + 2), 1ull);
                skip = 1;
            }
            s_skip_writes = skip;
            if (tile == P.n_tiles - 1) {
                P.totals[0] = static_cast<int64_t>(pre_l + my_lines);
                P.totals[1] = static_cast<int64_t>(pre_s + my_spans);
            }
        }
        __syncthreads();
        const int64_t line_base = s_line_base, span_base = s_span_base;
        const bool skip_writes = s_skip_writes != 0;

        // ---- E: capture automaton for the matched lines (spans straight to their final place)
        if (!skip_writes && !P.cap.match_only) {
            for (uint32_t i = threadIdx.x; i < n_t; i += kT) {
                const int32_t e = s_ext[i];
                if (e < 0) continue;
                const FastExtDev fx = P.cap_fast.ext[e];
                const ExtDev x = P.cap.ext[e];
                const uint32_t tab_abs = cls_abs + fx.tab_off;
                const uint32_t rel = s_start[i];
                uint32_t q = rel & ~7u;
                const uint32_t lo = rel & 7u;
                uint32_t st = lo ? (fx.n_states + lo - 1) * fx.row_bytes : 0u;
                uint32_t pos = q - rel;  // wraps while skipping: stored to the dummy register only
                for (;;) {
                    const uint4 v = q < static_cast<uint32_t>(kBuf) ? lds128(text_abs + q * 2) : load_chunk_global(P.text, t0 + q, P.n_units);
                    const uint32_t st0 = st;
                    bool slow = ((v.x | v.y | v.z | v.w) & 0xFF80FF80u) != 0u;
                    if (!slow) {
                        cap_step<0>(st, v.x, cls_abs, tab_abs, reg_abs, pos);
                        cap_step<2>(st, v.x, cls_abs, tab_abs, reg_abs, pos + 1);
                        cap_step<0>(st, v.y, cls_abs, tab_abs, reg_abs, pos + 2);
                        cap_step<2>(st, v.y, cls_abs, tab_abs, reg_abs, pos + 3);
                        cap_step<0>(st, v.z, cls_abs, tab_abs, reg_abs, pos + 4);
                        cap_step<2>(st, v.z, cls_abs, tab_abs, reg_abs, pos + 5);
                        cap_step<0>(st, v.w, cls_abs, tab_abs, reg_abs, pos + 6);
                        cap_step<2>(st, v.w, cls_abs, tab_abs, reg_abs, pos + 7);
                        slow = st == fx.slow_off;
                    }
                    if (slow)
                        st = tdfa_slow_chunk(P.cap, x, fx, st0, P.text, t0 + q, t0 + rel, P.n_units, reg_abs, reg_stride, len_off);
                    if (st >= fx.dead_off) break;
                    q += 8;
                    pos += 8;
                }
                int32_t* out = P.spans + span_base + s_spoff[i];
                bool ok = st >= fx.frz_off;
                uint32_t s = 0;
                if (ok) {
                    s = (st - fx.frz_off) / fx.row_bytes;
                    ok = __ldg(P.cap.tdfa_accepting + x.acc_off + s) != 0;
                }
                if (!ok) {
                    s_ext[i] = -2 - e;
                    for (uint32_t k = 0; k < x.n_slots; ++k) out[k] = -1;
                    continue;
                }
                const uint8_t* __restrict__ fin = P.cap.tdfa_fin + x.fin_off + s * x.n_slots;
                const int32_t len = static_cast<int32_t>(lds32(reg_abs + len_off));
                for (uint32_t k = 0; k < x.n_slots; ++k) {
                    const uint32_t r = __ldg(fin + k);
                    out[k] = r == 0xFFu ? -1 : (r == 0xFEu ? len : static_cast<int32_t>(lds32(reg_abs + r * reg_stride)));
                }
            }
        }
        __syncthreads();

        // ---- F: per-line result rows + histogram
        if (!skip_writes) {
            for (uint32_t i = threadIdx.x; i < n_t; i += kT) {
                const int32_t e = s_ext[i];
                P.ext_id[line_base + i] = e;
                P.line_off[line_base + i] = t0 + s_start[i];
                P.span_off[line_base + i] = span_base + s_spoff[i];
                const uint32_t bin = e >= 0 ? static_cast<uint32_t>(e) : (e == -1 ? P.n_ext : P.n_ext + 1);
                if (smem_hist) atomicAdd(&s_hist[bin], 1u);
                else atomicAdd(P.hist + bin, 1ull);
            }
            if (tile == P.n_tiles - 1 && threadIdx.x == 0) {
                const int64_t nl = line_base + n_t;
                // line i spans [line_off[i], line_off[i+1] - 1): a text that does not end in '\n' gets n_units + 1
                P.line_off[nl] = P.n_units + (P.text[P.n_units - 1] == 0x0A ? 0 : 1);
                P.span_off[nl] = span_base + span_t;
            }
            __syncthreads();
            if (smem_hist)
                for (uint32_t i = threadIdx.x; i < n_bins; i += kT)
                    if (s_hist[i]) atomicAdd(P.hist + i, static_cast<unsigned long long>(s_hist[i]));
        }
        __syncthreads();
    }
}

}  // namespace

size_t fused_smem_bytes(const FusedParams& P) {
    return static_cast<size_t>(P.dfa.n_rows) * 512 + static_cast<size_t>(P.cap_fast.image_words) * 4 +
           static_cast<size_t>(P.cap_fast.n_regs + 2) * kT * 4 + static_cast<size_t>(kBuf + 8) * 2 +
           static_cast<size_t>(kMaxLines) * (2 + 4 + 4) + 128;
}

bool k0_fused_supported(const FusedParams& P) {
    return P.dfa.enabled && (P.cap_fast.enabled || P.cap.match_only) && fused_smem_bytes(P) <= 110 * 1024;
}

void k0_fused_extract(const Launch& L, const FusedParams& P) {
    const size_t smem = fused_smem_bytes(P);
    cudaFuncSetAttribute(fused_extract_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fused_extract_kernel, kT, smem);
    if (per_sm < 1) per_sm = 1;
    int64_t cap = static_cast<int64_t>(L.sm_count) * per_sm;
    int g = static_cast<int>(P.n_tiles < cap ? P.n_tiles : cap);
    if (g < 1) g = 1;
    fused_extract_kernel<<<g, kT, smem, L.stream>>>(P);
}

}  // namespace gorp
