// K0' — one-pass persistent extraction kernel (sm_100a): every UTF-16 unit is looked at ONCE.
//
// The combined DFA (PolyMatcher.match) and the capture automaton of the winning extraction (Matcher.matches() +
// group(i)) are folded on the host into one automaton (host/fused.hpp); this kernel runs it:
//
//   per unit :  ent = LDS[row(ent) + 4*unit]              one shared-memory lookup, the only dependent chain
//               STS slot(ent)[thread] = position          "last position at which command list `slot` fired"
//   per line :  the '\n' column leads to the absorbing row of the line's OUTCOME (MISS | MATCH e | CAPTURE_FAIL e);
//               a group boundary = max over the (<= 4) op slots that write its register.
//
// Work decomposition: persistent CTAs take tiles of `tile_units` units by in-order ticket. A tile is pulled into
// shared memory by ONE TMA bulk copy (cp.async.bulk + mbarrier complete_tx), double-buffered: the copy of the next
// tile runs while the current one is processed. A CTA owns the lines that START in its tile (after a '\n' of the
// tile; tile 0 also owns offset 0): phase A finds them (128-bit LDS, __vcmpeq2, block scan), phase B walks one line
// per thread, phase C scans the span counts, phase D publishes/looks back the (lines, spans) prefix across tiles
// (decoupled look-back, one warp, 32 predecessors per probe), phase E writes the result rows. The host sizes the
// tile so that it holds slightly fewer lines than the CTA has threads; a tile with more line starts than threads
// raises FLAG_FALLBACK and the host reruns the batch with a smaller tile.
#include "device_common.cuh"

namespace gorp {

namespace {

using namespace dev;

constexpr unsigned long long kStAgg = 1ull << 62, kStPre = 2ull << 62;
constexpr uint32_t kOver = kOnePassOverhang;

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
__device__ __forceinline__ unsigned long long ld_acquire(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// 8 units at text position `pos` straight from global memory; units at or beyond n_units read as '\n'
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

// Entry layout in shared memory (rewritten from the raw table when the CTA starts):
//   bits 31..18  next row, in units of 16 bytes from the start of the row area
//   bits 17..14  zero
//   bits 13..0   op slot, in units of 4 bytes from the start of the slot area (slot id * blockDim)
// so that  next lookup address = (ent >> 14) + (rows_abs + 4*unit)   is a single LEA.HI on the dependent chain.
template <int kByte>
__device__ __forceinline__ void one_step(uint32_t& ent, uint32_t w, uint32_t rows_abs, uint32_t slot_abs, uint32_t pos) {
    uint32_t b, a;
    asm("prmt.b32 %0, %1, 0, %2;" : "=r"(b) : "r"(w), "n"(kByte == 0 ? 0x4440 : 0x4442));
    asm("mad.lo.u32 %0, %1, 4, %2;" : "=r"(a) : "r"(b), "r"(rows_abs));
    ent = lds32((ent >> 14) + a);
    uint32_t m, sa;
    asm("and.b32 %0, %1, 0x3FFF;" : "=r"(m) : "r"(ent));
    asm("mad.lo.u32 %0, %1, 4, %2;" : "=r"(sa) : "r"(m), "r"(slot_abs));
    sts32(sa, pos);
}

// A chunk that holds a unit >= 0x80: unit by unit through the column map (global, L1/L2 resident). A high surrogate
// followed by a low surrogate takes the PAIR column (java.util.regex consumes the pair as one character).
__device__ __noinline__ uint32_t slow_chunk(const OnePassDev& a, uint32_t ent, uint4 v, const uint16_t* __restrict__ text,
                                            int64_t q, int64_t n_units, uint32_t rows_abs, uint32_t slot_abs, uint32_t pos,
                                            uint32_t fin_ent) {
#pragma unroll 1
    for (int k = 0; k < 8; ++k) {
        if (ent >= fin_ent) break;
        const uint32_t u = unit_at(v, k);
        uint32_t col = u;
        if (u >= 0x80u) {
            col = __ldg(a.xcol + u);
            if ((u & 0xFC00u) == 0xD800u) {
                const int64_t p1 = q + k + 1;
                const uint32_t nx = k < 7 ? unit_at(v, k + 1) : (p1 < n_units ? __ldg(text + p1) : 0x0Au);
                if ((nx & 0xFC00u) == 0xDC00u) col = __ldg(a.pair_col + col);
            }
        }
        ent = lds32((ent >> 14) + rows_abs + col * 4);
        sts32(slot_abs + ((ent & 0x3FFFu) << 2), pos + k);
    }
    return ent;
}

// bit k (k < 4) = unit k of the pair of words (a, b) is '\n' (0x000A). Low and high bytes are gathered with PRMT, a
// unit is '\n' iff (low ^ 0x0A) | high == 0; exact zero-byte test, then the four flag bits (7, 15, 23, 31) are
// compressed with one multiply.
__device__ __forceinline__ uint32_t nl_bits4(uint32_t a, uint32_t b) {
    const uint32_t lo = __byte_perm(a, b, 0x6420), hi = __byte_perm(a, b, 0x7531);
    const uint32_t t = (lo ^ 0x0A0A0A0Au) | hi;
    const uint32_t z = ~(((t & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | t | 0x7F7F7F7Fu);  // 0x80 in every zero byte of t
    return (((z >> 7) * 0x00204081u) >> 21) & 0xFu;
}

// ---- named barriers: the CTA is kT worker threads (8 warps) + one look-back warp
constexpr uint32_t kIoThreads = 32;
constexpr uint32_t kBarWorkers = 1;  // workers only
constexpr uint32_t kBarCount = 2;    // workers arrive, look-back warp waits: the tile's line count is known (after phase A)
constexpr uint32_t kBarBase = 3;     // look-back warp arrives, workers wait: the tile's first result row is known
constexpr uint32_t kSortBins = 64;

__device__ __forceinline__ void bar_sync(uint32_t id, uint32_t n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(uint32_t id, uint32_t n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// exclusive scan over the kT worker threads (worker barrier inside)
__device__ __forceinline__ uint32_t block_scan(uint32_t v, uint32_t* s_warp, uint32_t* total, uint32_t kT) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = kT >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    bar_sync(kBarWorkers, kT);  // s_warp reuse
    if (lane == 31) s_warp[warp] = incl;
    bar_sync(kBarWorkers, kT);
    uint32_t base = 0, tot = 0;
    for (int w = 0; w < nw; ++w) {
        const uint32_t x = s_warp[w];
        if (w < warp) base += x;
        tot += x;
    }
    *total = tot;
    return base + incl - v;
}

__global__ void __launch_bounds__(512) onepass_kernel(OnePassParams P) {
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t kT = blockDim.x - kIoThreads;  // worker threads
    const uint32_t kAll = blockDim.x;
    const OnePassDev& A = P.a;
    const uint32_t T = P.tile_units, kBuf = T + kOver;
    // ---- carve shared memory (every area 16-byte aligned)
    uint32_t* s_rows = reinterpret_cast<uint32_t*>(smem);
    uint32_t* s_res = s_rows + A.n_rows * A.width;
    int32_t* s_oext = reinterpret_cast<int32_t*>(s_res + ((A.n_outcomes * A.max_slots + 3) & ~3u));
    uint32_t* s_ocnt = reinterpret_cast<uint32_t*>(s_oext + ((A.n_outcomes + 3) & ~3u));  // valid span entries per outcome
    uint32_t* s_slots = s_ocnt + ((A.n_outcomes + 3) & ~3u);
    uint16_t* s_text0 = reinterpret_cast<uint16_t*>(s_slots + A.n_slots * kT);
    uint16_t* s_text1 = s_text0 + kBuf + 8;
    uint16_t* s_start = s_text1 + kBuf + 8;  // [kT] line starts of the tile, in text order
    uint16_t* s_perm = s_start + kT;         // [kT] worker thread -> line of the tile (sorted by length)
    __shared__ __align__(8) unsigned long long s_bar[2];
    __shared__ uint32_t s_warp[16];
    __shared__ uint32_t s_bins[kSortBins];
    __shared__ uint32_t s_hist[kOnePassHistBins];
    __shared__ long long s_next_tile;
    __shared__ long long s_cnt_tile, s_line_base;  // workers -> look-back warp: tile id (-1 = done); and back: first row
    __shared__ uint32_t s_cnt_lines, s_cnt_dense;
    __shared__ int s_skip_writes;

    const uint32_t rows_abs = static_cast<uint32_t>(__cvta_generic_to_shared(s_rows));
    const uint32_t text_abs0 = static_cast<uint32_t>(__cvta_generic_to_shared(s_text0));
    const uint32_t text_abs1 = static_cast<uint32_t>(__cvta_generic_to_shared(s_text1));
    const uint32_t bar0 = static_cast<uint32_t>(__cvta_generic_to_shared(&s_bar[0]));

    const uint32_t row_q = A.width / 4;  // row size in 16-byte units
    for (uint32_t i = threadIdx.x; i < A.n_rows * A.width; i += kAll) {
        const uint32_t raw = __ldg(A.rows + i);
        s_rows[i] = (((raw >> 16) * row_q) << 18) | ((raw & 0xFFFFu) * kT);
    }
    for (uint32_t i = threadIdx.x; i < A.n_outcomes * A.max_slots; i += kAll) s_res[i] = __ldg(A.out_res + i);
    for (uint32_t i = threadIdx.x; i < A.n_outcomes; i += kAll) {
        const int32_t e = __ldg(A.out_ext + i);
        s_oext[i] = e;
        s_ocnt[i] = e >= 0 ? __ldg(P.slots_per_ext + e) : 0u;
    }
    __shared__ uint32_t s_init[16];
    for (uint32_t i = threadIdx.x; i < A.n_init && i < 16; i += kAll) s_init[i] = __ldg(A.init_slots + i);
    const uint32_t n_bins = P.n_ext + 2;
    const bool smem_hist = n_bins <= kOnePassHistBins;
    for (uint32_t i = threadIdx.x; i < kOnePassHistBins; i += kAll) s_hist[i] = 0;

    auto issue_load = [&](int64_t tile, uint32_t b) {  // worker thread 0 only
        const int64_t t0 = tile * T;
        const int64_t avail = P.n_units - t0 < kBuf + 8 ? P.n_units - t0 : kBuf + 8;
        const uint32_t bulk_units = static_cast<uint32_t>(avail) & ~7u;
        if (bulk_units) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(bar0 + 8 * b, bulk_units * 2);
            tma_bulk_g2s(b ? text_abs1 : text_abs0, P.text + t0, bulk_units * 2, bar0 + 8 * b);
        }
    };

    if (threadIdx.x == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const long long first = atomicAdd(P.ticket, 1u);
        s_next_tile = first;
        if (first < P.n_tiles) issue_load(first, 0);
    }
    __syncthreads();

    if (threadIdx.x >= kT) {
        // =============================================================== look-back warp
        // As soon as the workers know how many lines START in the tile (phase A), publish that count, look back over
        // the predecessor tiles (decoupled look-back, 32 tiles per probe) and hand the tile's first result row to
        // the workers, who are walking the tile's lines meanwhile. Predecessors publish at the beginning of THEIR
        // tile, so this never waits for a predecessor's walk.
        const uint32_t lane = threadIdx.x - kT;
        for (;;) {
            bar_sync(kBarCount, kAll);
            const long long tile = s_cnt_tile;
            if (tile < 0) break;
            const unsigned long long my_lines = s_cnt_lines;
            const bool too_dense = s_cnt_dense != 0;
            unsigned long long pre_l = 0;
            if (tile > 0) {
                if (lane == 0) st_release(P.tile_status + tile, kStAgg | my_lines);
                for (int64_t j = tile - 1;; j -= 32) {
                    const int64_t idx = j - lane;
                    unsigned long long v = kStPre;  // before tile 0: an empty inclusive prefix
                    if (idx >= 0) {
                        v = ld_acquire(P.tile_status + idx);
                        while ((v >> 62) == 0) {
                            __nanosleep(32);
                            v = ld_acquire(P.tile_status + idx);
                        }
                    }
                    const uint32_t pmask = __ballot_sync(0xffffffffu, (v >> 62) == 2);
                    const uint32_t first = pmask ? static_cast<uint32_t>(__ffs(pmask)) - 1u : 32u;
                    unsigned long long l = lane <= first ? (v & ~(3ull << 62)) : 0ull;  // aggregates, then one inclusive prefix
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
                    pre_l += l;
                    if (pmask) break;
                }
            }
            const long long line_base = static_cast<long long>(pre_l), line_end = line_base + static_cast<long long>(my_lines);
            if (lane == 0) {
                st_release(P.tile_status + tile, kStPre | static_cast<unsigned long long>(line_end));
                const bool over = line_end > P.cap_lines;
                if (too_dense) atomicOr(reinterpret_cast<unsigned long long*>(P.totals + 2), 2ull);
                if (over) atomicOr(reinterpret_cast<unsigned long long*>(P.totals + 2), 1ull);
                if (tile == P.n_tiles - 1) {
                    P.totals[0] = line_end;
                    // line i spans [line_off[i], line_off[i+1] - 1): a text that does not end in '\n' gets n_units + 1
                    if (!over) P.line_off[line_end] = P.n_units + (P.text[P.n_units - 1] == 0x0A ? 0 : 1);
                }
                s_line_base = line_base;
                s_skip_writes = (over || too_dense) ? 1 : 0;
            }
            __syncwarp();
            __threadfence_block();
            bar_arrive(kBarBase, kAll);
        }
        return;
    }

    // =================================================================== workers
    uint32_t slot_abs = static_cast<uint32_t>(__cvta_generic_to_shared(s_slots)) + threadIdx.x * 4;
    asm volatile("mov.u32 %0, %0;" : "+r"(slot_abs));  // one opaque register: keeps the per-step store address at LOP3 + LEA
    const uint32_t slot_stride = kT * 4;
    const uint32_t fin_ent = (A.fin_base * row_q) << 18;
    const uint32_t inv_row_q = 65536u / row_q + 1u;
    const uint32_t stride = P.span_stride;
    int64_t tile = s_next_tile;
    uint32_t buf = 0, phase = 0;  // phase bit b = parity to wait for on barrier b
    // thread 0 takes the ticket of the tile after the next one early (its latency hides behind the result rows)
    long long ticket_ahead = threadIdx.x == 0 && tile < P.n_tiles ? static_cast<long long>(atomicAdd(P.ticket, 1u)) : 0;
    // optional per-phase cycle accounting (GORP_ONEPASS_DEBUG=1): thread 0 and the first thread of the last worker warp
    const bool dbg = P.debug != nullptr && (threadIdx.x == 0 || threadIdx.x == kT - 32);
    long long dbg_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, dbg_prev = dbg ? clock64() : 0;
#define GORP_PH(i)                         \
    if (dbg) {                             \
        const long long t_ = clock64();    \
        dbg_acc[i] += t_ - dbg_prev;       \
        dbg_prev = t_;                     \
    }

    while (tile < P.n_tiles) {
        bar_sync(kBarWorkers, kT);  // everyone has read s_next_tile; the other text buffer is no longer being read
        GORP_PH(0)
        if (threadIdx.x == 0) {
            const long long nx = ticket_ahead;
            s_next_tile = nx;
            if (nx < P.n_tiles) issue_load(nx, buf ^ 1);
        }
        const int64_t t0 = tile * T;
        const uint32_t text_abs = buf ? text_abs1 : text_abs0;
        uint16_t* s_text = buf ? s_text1 : s_text0;
        const int64_t avail = P.n_units - t0 < kBuf + 8 ? P.n_units - t0 : kBuf + 8;  // > 0
        const uint32_t bulk_units = static_cast<uint32_t>(avail) & ~7u;  // == kBuf + 8 except at the end of the text
        for (uint32_t i = bulk_units + threadIdx.x; i < kBuf + 8; i += kT)
            s_text[i] = i < avail ? __ldg(P.text + t0 + i) : static_cast<uint16_t>(0x0A);
        if (threadIdx.x < kSortBins) s_bins[threadIdx.x] = 0;
        if (bulk_units) {
            mbar_wait(bar0 + 8 * buf, (phase >> buf) & 1u);
            phase ^= 1u << buf;
        }
        GORP_PH(1)
        bar_sync(kBarWorkers, kT);
        GORP_PH(2)

        // ---- A: line starts owned by this tile = (position of a '\n' in [t0, t0+T)) + 1, if < n_units
        const uint32_t per = P.per;  // units per thread, multiple of 8, <= 64
        unsigned long long nl_mask = 0;
        {
            const uint32_t u0 = threadIdx.x * per;
            for (uint32_t j = 0; j < per / 8; ++j) {
                const uint4 v = lds128(text_abs + (u0 + j * 8) * 2);
                nl_mask |= static_cast<unsigned long long>(nl_bits4(v.x, v.y) | (nl_bits4(v.z, v.w) << 4)) << (j * 8);
            }
            // a '\n' at position p starts a line only if p + 1 < n_units (the padding beyond the text is '\n' too)
            const int64_t room = P.n_units - 1 - (t0 + u0);
            if (room < static_cast<int64_t>(per)) nl_mask = room <= 0 ? 0ull : (nl_mask & ((1ull << room) - 1ull));
        }
        GORP_PH(3)
        const uint32_t extra = tile == 0 ? 1u : 0u;  // the line at offset 0
        uint32_t n_t;
        uint32_t my = block_scan(static_cast<uint32_t>(__popcll(nl_mask)), s_warp, &n_t, kT) + extra;
        n_t += extra;
        const bool too_dense = n_t > kT;
        if (threadIdx.x == 0) {
            s_cnt_tile = tile;
            s_cnt_lines = too_dense ? 0u : n_t;
            s_cnt_dense = too_dense ? 1u : 0u;
        }
        __threadfence_block();
        bar_arrive(kBarCount, kAll);  // the look-back for this tile runs while the lines are walked
        if (!too_dense) {
            if (extra && threadIdx.x == 0) s_start[0] = 0;
            const uint32_t u0 = threadIdx.x * per;
            while (nl_mask) {
                const int k = __ffsll(static_cast<long long>(nl_mask)) - 1;
                nl_mask &= nl_mask - 1;
                s_start[my++] = static_cast<uint16_t>(u0 + k + 1);  // 1 .. T
            }
        }
        bar_sync(kBarWorkers, kT);
        GORP_PH(4)

        // ---- A': counting sort of the tile's lines by the number of 16-byte chunks they span, so that the 32 lines
        // of a warp take (almost) the same number of loop iterations. The tile's last line ends beyond the tile:
        // top bin.
        const bool have_line = !too_dense && threadIdx.x < n_t;
        uint32_t key = 0, rank = 0;
        if (have_line) {
            const uint32_t a = s_start[threadIdx.x];
            key = kSortBins - 1;
            if (threadIdx.x + 1 < n_t) {
                const uint32_t chunks = ((a & 7u) + (s_start[threadIdx.x + 1] - a) + 7u) >> 3;
                key = chunks < kSortBins - 1 ? chunks : kSortBins - 1;
            }
            rank = atomicAdd(&s_bins[key], 1u);
        }
        bar_sync(kBarWorkers, kT);
        {   // exclusive prefix over the 64 bins, redundantly per warp: lane l owns bins 2l and 2l+1
            const uint32_t lane = threadIdx.x & 31u;
            const uint32_t b0 = s_bins[2 * lane], b1 = s_bins[2 * lane + 1];
            uint32_t incl = b0 + b1;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            const uint32_t excl = incl - (b0 + b1);
            const uint32_t e = __shfl_sync(0xffffffffu, excl, key >> 1), f = __shfl_sync(0xffffffffu, b0, key >> 1);
            if (have_line) s_perm[rank + e + ((key & 1u) ? f : 0u)] = static_cast<uint16_t>(threadIdx.x);
        }
        bar_sync(kBarWorkers, kT);
        GORP_PH(5)

        // ---- B: one line per thread through the one-pass automaton
        uint32_t outcome = 0, line = 0, rel = 0;
        if (have_line) {
            line = s_perm[threadIdx.x];
            for (uint32_t k = 0; k < A.n_init; ++k)
                sts32(slot_abs + (k < 16 ? s_init[k] : __ldg(A.init_slots + k)) * slot_stride, 0xFFFFFFFFu);
            rel = s_start[line];
            uint32_t q = rel & ~7u;
            const uint32_t lo = rel & 7u;
            uint32_t ent = lo ? ((A.skip_base + lo - 1) * row_q) << 18 : 0u;
            uint32_t pos = q - rel;  // wraps while skipping: only ever stored to the dummy slot
            do {
                const uint4 v = q < kBuf ? lds128(text_abs + q * 2) : load_chunk_global(P.text, t0 + q, P.n_units);
                if (((v.x | v.y | v.z | v.w) & 0xFF80FF80u) == 0u) {
                    one_step<0>(ent, v.x, rows_abs, slot_abs, pos);
                    one_step<2>(ent, v.x, rows_abs, slot_abs, pos + 1);
                    one_step<0>(ent, v.y, rows_abs, slot_abs, pos + 2);
                    one_step<2>(ent, v.y, rows_abs, slot_abs, pos + 3);
                    one_step<0>(ent, v.z, rows_abs, slot_abs, pos + 4);
                    one_step<2>(ent, v.z, rows_abs, slot_abs, pos + 5);
                    one_step<0>(ent, v.w, rows_abs, slot_abs, pos + 6);
                    one_step<2>(ent, v.w, rows_abs, slot_abs, pos + 7);
                } else {
                    ent = slow_chunk(A, ent, v, P.text, t0 + q, P.n_units, rows_abs, slot_abs, pos, fin_ent);
                }
                q += 8;
                pos += 8;
            } while (ent < fin_ent);
            outcome = (((ent >> 18) - A.fin_base * row_q) * inv_row_q) >> 16;  // exact: a multiple of row_q below 2^14
        }

        // ---- C: result rows. Row = (first row of the tile, from the look-back warp) + line index inside the tile.
        // Once every walk is over this tile's text buffer is dead: the rows are staged in it in line order and, when
        // the look-back warp has delivered the first row, copied out with coalesced (vector) stores. Rows too wide
        // to be staged go out directly.
        GORP_PH(6)
        const bool staged = (kT * 4u * (1u + stride)) <= (kBuf + 8u) * 2u;
        int32_t* st_ext = reinterpret_cast<int32_t*>(s_text);
        int32_t* st_spans = st_ext + kT;
        if (staged) bar_sync(kBarWorkers, kT);
        else bar_sync(kBarBase, kAll);
        if (threadIdx.x == 0 && s_next_tile < P.n_tiles) ticket_ahead = static_cast<long long>(atomicAdd(P.ticket, 1u));
        if (dbg) dbg_acc[11] += st_ext[0] & 1;  // consume a value: the clock below is read after the barrier
        GORP_PH(7)
        int32_t ext = -1;
        if (have_line) {
            ext = s_oext[outcome];
            const uint32_t cnt = s_ocnt[outcome];
            const uint32_t* res = s_res + outcome * A.max_slots;
            const uint32_t* my_slots = s_slots + threadIdx.x;
            int32_t* out = staged ? st_spans + line * stride : P.spans + (s_line_base + line) * stride;
            if (staged || !s_skip_writes) {
                for (uint32_t k = 0; k < stride; ++k) {
                    int32_t val = -1;
                    for (uint32_t packed = k < cnt ? res[k] : 0u; packed; packed >>= 8)
                        val = max(val, static_cast<int32_t>(my_slots[(packed & 0xFFu) * kT]));
                    out[k] = val;
                }
                if (staged) {
                    st_ext[line] = ext;
                } else {
                    P.ext_id[s_line_base + line] = ext;
                    P.line_off[s_line_base + line] = t0 + rel;
                }
            }
        }
        // histogram: one shared-memory atomic per distinct outcome of the warp
        if (smem_hist) {
            const uint32_t bin = !have_line ? 0xFFFFFFFFu : ext >= 0 ? static_cast<uint32_t>(ext) : (ext == -1 ? P.n_ext : P.n_ext + 1);
            const uint32_t peers = __match_any_sync(0xffffffffu, bin);
            if (have_line && (threadIdx.x & 31u) == static_cast<uint32_t>(__ffs(peers)) - 1u) atomicAdd(&s_hist[bin], static_cast<uint32_t>(__popc(peers)));
        } else if (have_line) {
            atomicAdd(P.hist + (ext >= 0 ? static_cast<uint32_t>(ext) : (ext == -1 ? P.n_ext : P.n_ext + 1)), 1ull);
        }
        GORP_PH(8)
        if (staged) {
            bar_sync(kBarBase, kAll);
            const bool skip_writes = s_skip_writes != 0;
            const int64_t line_base = s_line_base;
            if (dbg) dbg_acc[11] += line_base & 1;
            GORP_PH(9)
            if (!skip_writes && !too_dense) {
                if (threadIdx.x < n_t) {
                    P.ext_id[line_base + threadIdx.x] = st_ext[threadIdx.x];
                    P.line_off[line_base + threadIdx.x] = t0 + s_start[threadIdx.x];
                }
                const uint32_t total = n_t * stride;
                int32_t* dst = P.spans + line_base * stride;
                if (((line_base * stride) & 3) == 0) {
                    const uint32_t n4 = total >> 2;
                    for (uint32_t i = threadIdx.x; i < n4; i += kT) reinterpret_cast<int4*>(dst)[i] = reinterpret_cast<const int4*>(st_spans)[i];
                    for (uint32_t i = (n4 << 2) + threadIdx.x; i < total; i += kT) dst[i] = st_spans[i];
                } else {
                    for (uint32_t i = threadIdx.x; i < total; i += kT) dst[i] = st_spans[i];
                }
            }
        }
        GORP_PH(10)
        dbg_acc[11] += 4;
        tile = s_next_tile;
        buf ^= 1;
    }
    if (dbg)
        for (int i = 0; i < 12; ++i) P.debug[(blockIdx.x * 2 + (threadIdx.x ? 1 : 0)) * 12 + i] = dbg_acc[i];
#undef GORP_PH
    if (threadIdx.x == 0) s_cnt_tile = -1;
    __threadfence_block();
    bar_arrive(kBarCount, kAll);
    if (smem_hist) {
        bar_sync(kBarWorkers, kT);
        for (uint32_t i = threadIdx.x; i < n_bins; i += kT)
            if (s_hist[i]) atomicAdd(P.hist + i, static_cast<unsigned long long>(s_hist[i]));
    }
}

}  // namespace

size_t onepass_smem_bytes(const OnePassDev& a, uint32_t threads, uint32_t tile_units) {
    size_t b = static_cast<size_t>(a.n_rows) * a.width * 4;
    b += static_cast<size_t>((a.n_outcomes * a.max_slots + 3) & ~3u) * 4;
    b += 2 * static_cast<size_t>((a.n_outcomes + 3) & ~3u) * 4;
    b += static_cast<size_t>(a.n_slots) * threads * 4;
    b += 2 * static_cast<size_t>(tile_units + kOnePassOverhang + 8) * 2;
    b += static_cast<size_t>(threads) * (2 + 2);  // s_start, s_perm
    return b + 128;
}

bool k0_onepass_plan(const OnePassDev& a, double lines_per_unit, uint32_t shrink, uint32_t* threads, uint32_t* tile_units) {
    if (!a.enabled) return false;
    // `threads` worker threads walk one line each; the tile (threads * per units) should hold at most ~0.9 * threads
    // line starts. per is a multiple of 8 and <= 64 (one 64-bit newline mask per thread).
    const uint32_t kT = 256;
    if (static_cast<uint64_t>(a.n_slots) * kT > 0x3FFFu) return false;
    const double lpu = lines_per_unit > 1e-9 ? lines_per_unit : 1e-9;
    uint32_t per = 8;
    for (uint32_t cand = 64; cand >= 8; cand -= 8)
        if (cand * lpu <= 0.90) {
            per = cand;
            break;
        }
    for (uint32_t s = 0; s < shrink && per > 8; ++s) per = per > 16 ? (per * 3 / 4) & ~7u : 8;
    for (;;) {
        if (onepass_smem_bytes(a, kT, kT * per) <= 113 * 1024) break;  // two CTAs per SM
        if (per == 8) {
            if (onepass_smem_bytes(a, kT, kT * per) <= 226 * 1024) break;
            return false;
        }
        per -= 8;
    }
    *threads = kT;
    *tile_units = kT * per;
    return true;
}

int k0_onepass_grid(const Launch& L, const OnePassParams& P, uint32_t threads) {
    const size_t smem = onepass_smem_bytes(P.a, threads, P.tile_units);
    const int block = static_cast<int>(threads + 32);  // + the look-back warp
    allow_max_dynamic_smem(onepass_kernel);
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, onepass_kernel, block, smem);
    if (per_sm < 1) per_sm = 1;
    const int64_t cap = static_cast<int64_t>(L.sm_count) * per_sm;
    int g = static_cast<int>(P.n_tiles < cap ? P.n_tiles : cap);
    return g < 1 ? 1 : g;
}

void k0_onepass_extract(const Launch& L, const OnePassParams& P, uint32_t threads) {
    const size_t smem = onepass_smem_bytes(P.a, threads, P.tile_units);
    const int g = k0_onepass_grid(L, P, threads);
    onepass_kernel<<<g, static_cast<int>(threads + 32), smem, L.stream>>>(P);
}

}  // namespace gorp
