// K0c — one-pass "chunk owner" extraction kernel (sm_100a): every UTF-16 unit is looked at once, no shared-memory
// staging of the text, no barrier inside the walk.
//
// The automaton is the one of host/fused.hpp (combined DFA x capture automata folded into one table):
//
//   per unit :  ent = LDS[row(ent) + 4*unit]              one shared-memory lookup, the only dependent chain
//               STS slot(ent)[thread] = position          "last position at which command list `slot` fired"
//   per line :  the '\n' column leads to the absorbing row of the line's OUTCOME (MISS | MATCH e | CAPTURE_FAIL e)
//
// Work decomposition: the text is cut into fixed chunks of kChunkUnits units, one per thread; a CTA takes a tile of
// blockDim.x consecutive chunks by in-order ticket. A thread owns the lines whose PRECEDING '\n' lies in its chunk
// (thread 0 of tile 0 also owns the line at offset 0) and walks them one after the other straight from global
// memory with 256-bit loads (its chunk is a private sequential stream: one 32-byte sector per load, so nothing
// depends on L1 retention), running past the end of its chunk to finish its last line. All threads therefore walk about
// kChunkUnits units whatever the line lengths are, so the lanes of a warp stay busy without any sorting, and the
// only block-wide steps are one scan of the per-chunk line counts at the start of the tile (newline pre-scan, done
// per warp with coalesced loads: iteration i of a warp covers exactly the chunk of lane i) and the decoupled
// look-back of the tile's first result row, done by warp 0 while the other warps are already walking. Result rows go
// out when a line ends, batched every 2 loop iterations so that the lanes of a warp that have a finished line write
// together; two banks of op slots per thread keep the finished line's positions alive meanwhile. Slots hold
// tile-relative positions and are cleared once per tile: a value below the line's own start is "not written", so no
// per-line initialisation is needed. Occupancy: the table + slots are the only shared memory (no text buffers), which is
// what lets 32 warps per SM hide the lookup latency — the measured reason this beats the TMA-staged tile kernel.
#include <cstdlib>

#include "device_common.cuh"

namespace gorp {

namespace {

using namespace dev;

// Entry layout in shared memory (rewritten from the raw table when the CTA starts):
//   bits 31..18  next row, in units of 16 bytes from the start of the row area
//   bits 13..0   op slot, in units of 4 bytes from the start of the slot bank (slot id * blockDim)
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

struct Pending {  // a finished line whose result row has not been written yet
    uint32_t outcome, idx, bank_abs;
    uint32_t start;  // tile-relative
};

// kDbg: GORP_ONEPASS_DEBUG build of the same kernel — thread-cycles per phase, summed per CTA into P.debug[16 b + 4..11]:
// pre-scan, block scan + barrier, clear + walk without the two below, waiting for the tile's first row (look-back),
// result rows, end-of-tile barrier wait, look-back (warp 0), tiles
// (A/B, round 2: variants with 384 / 256 threads per CTA that load the next 16-unit block while the current one is walked —
// 8 more registers — ran at 13.8 / 13.9 ms against 11.65 ms: the walk phase is not bound by the latency of its block loads.)
template <bool kDbg>
__global__ void __launch_bounds__(512, 2) chunkwalk_kernel(OnePassParams P) {
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t kT = blockDim.x;
    const OnePassDev& A = P.a;
    constexpr uint32_t C = kChunkUnits;
    // ---- carve shared memory (every area 16-byte aligned)
    uint32_t* s_rows = reinterpret_cast<uint32_t*>(smem);
    uint32_t* s_res = s_rows + A.n_rows * A.width;
    int32_t* s_oext = reinterpret_cast<int32_t*>(s_res + ((A.n_outcomes * A.max_slots + 3) & ~3u));
    uint32_t* s_obin = reinterpret_cast<uint32_t*>(s_oext + ((A.n_outcomes + 3) & ~3u));  // histogram bin per outcome
    uint32_t* s_slots = s_obin + ((A.n_outcomes + 3) & ~3u);  // [2 banks][n_slots][kT]
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_hist[kOnePassHistBins];
    __shared__ long long s_tile, s_base;
    __shared__ int s_skip_writes;
    __shared__ unsigned int s_flag;  // (tile + 1) once s_base / s_skip_writes of the tile are valid

    const uint32_t rows_abs = static_cast<uint32_t>(__cvta_generic_to_shared(s_rows));
    const uint32_t row_q = A.width / 4;  // row size in 16-byte units
    for (uint32_t i = threadIdx.x; i < A.n_rows * A.width; i += kT) {
        const uint32_t raw = __ldg(A.rows + i);
        s_rows[i] = (((raw >> 16) * row_q) << 18) | ((raw & 0xFFFFu) * kT);
    }
    // group boundary recipes: 0 = no writer; one writer: byte offset of its slot inside a bank (slot ids start at 1);
    // several writers: bit 31 | one slot id per byte (ids < 128 here: the plan keeps n_slots * blockDim below 2^14)
    for (uint32_t i = threadIdx.x; i < A.n_outcomes * A.max_slots; i += kT) {
        const uint32_t packed = __ldg(A.out_res + i);
        s_res[i] = packed < 256u ? packed * kT * 4u : 0x80000000u | packed;
    }
    for (uint32_t i = threadIdx.x; i < A.n_outcomes; i += kT) {
        const int32_t e = __ldg(A.out_ext + i);
        s_oext[i] = e;
        s_obin[i] = e >= 0 ? static_cast<uint32_t>(e) : (e == -1 ? P.n_ext : P.n_ext + 1);  // histogram bin of the outcome
    }
    const uint32_t n_bins = P.n_ext + 2;
    const bool smem_hist = n_bins <= kOnePassHistBins;
    for (uint32_t i = threadIdx.x; i < kOnePassHistBins; i += kT) s_hist[i] = 0;
    if (threadIdx.x == 0) s_flag = 0;

    const uint32_t slot_abs0 = static_cast<uint32_t>(__cvta_generic_to_shared(s_slots)) + threadIdx.x * 4;
    const uint32_t slot_stride = kT * 4, bank_bytes = A.n_slots * slot_stride;
    const uint32_t fin_ent = (A.fin_base * row_q) << 18;
    const uint32_t inv_row_q = 65536u / row_q + 1u;
    const uint32_t stride = P.span_stride;
    const uint32_t len_off = (A.n_slots - 1) * slot_stride;  // the LEN slot is the last one
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, n_warps = kT >> 5;
    bool base_known = false;
    int64_t row0 = 0;  // first result row of this thread in the current tile
    bool skip_writes = false;
    unsigned long long dc_pre = 0, dc_scan = 0, dc_walk = 0, dc_poll = 0, dc_flush = 0, dc_endwait = 0, dc_look = 0;
    long long dc_t = 0;
    auto dtick = [&](unsigned long long& acc) {
        if (kDbg) {
            const long long now = clock64();
            acc += static_cast<unsigned long long>(now - dc_t);
            dc_t = now;
        }
    };

    // writes the result row of a finished line
    auto flush = [&](const Pending& pd, int64_t tile, int64_t tile0) {
        if (kDbg) dtick(dc_walk);
        if (!base_known) {
            // the tile's first row is not known yet (look-back of warp 0 still under way). A/B (profiles/README.md round 2):
            // sleeping in this loop instead of polling changes nothing (11.56 vs 11.48 ms) — GORP_CW_PREFETCH=4 sleeps
            while (*reinterpret_cast<volatile unsigned int*>(&s_flag) != static_cast<unsigned int>(tile + 1)) {
                if (P.per & 4u) __nanosleep(64);
            }
            __threadfence_block();
            row0 += *reinterpret_cast<volatile long long*>(&s_base);
            skip_writes = *reinterpret_cast<volatile int*>(&s_skip_writes) != 0;
            base_known = true;
        }
        if (kDbg) dtick(dc_poll);
        const uint32_t bin = s_obin[pd.outcome];
        if (smem_hist) atomicAdd(&s_hist[bin], 1u);
        else atomicAdd(P.hist + bin, 1ull);
        if (skip_writes) return;
        const int64_t row = row0 + pd.idx;
        P.ext_id[row] = s_oext[pd.outcome];
        P.line_off[row] = tile0 + pd.start;
        const uint32_t* res = s_res + pd.outcome * A.max_slots;
        int32_t* out = P.spans + row * stride;
        const int32_t start = static_cast<int32_t>(pd.start);
        // a boundary = the position held by its writer slot; slots hold tile-relative positions, anything below the line's
        // start is a left-over of an earlier line (or the per-tile clear value): the group did not participate
        auto single = [&](uint32_t recipe) {  // recipe 0 reads the dummy slot: discarded
            const int32_t val = static_cast<int32_t>(lds32(pd.bank_abs + recipe));
            return (recipe == 0u || val < start) ? -1 : val - start;
        };
        auto value = [&](uint32_t recipe) {  // several writers: the latest of them
            if (!(recipe & 0x80000000u)) return single(recipe);
            recipe &= 0x7FFFFFFFu;
            int32_t val = static_cast<int32_t>(lds32(pd.bank_abs + (recipe & 0xFFu) * slot_stride));
            for (recipe >>= 8; recipe; recipe >>= 8)
                val = max(val, static_cast<int32_t>(lds32(pd.bank_abs + (recipe & 0xFFu) * slot_stride)));
            return val < start ? -1 : val - start;
        };
        if ((stride & 7u) == 0) {  // 32-byte aligned rows: one 256-bit streaming store per 8 entries
            for (uint32_t k = 0; k < stride; k += 8) {
                const uint4 r4 = *reinterpret_cast<const uint4*>(res + k), r5 = *reinterpret_cast<const uint4*>(res + k + 4);
                // every entry through the single-writer path first (a several-writer recipe reads the dummy slot there),
                // then only the several-writer entries again, one by one
                auto fast = [&](uint32_t recipe) { return single(recipe & 0x80000000u ? 0u : recipe); };
                int32_t v0 = fast(r4.x), v1 = fast(r4.y), v2 = fast(r4.z), v3 = fast(r4.w);
                int32_t v4 = fast(r5.x), v5 = fast(r5.y), v6 = fast(r5.z), v7 = fast(r5.w);
                if ((r4.x | r4.y | r4.z | r4.w | r5.x | r5.y | r5.z | r5.w) & 0x80000000u) {
                    if (r4.x & 0x80000000u) v0 = value(r4.x);
                    if (r4.y & 0x80000000u) v1 = value(r4.y);
                    if (r4.z & 0x80000000u) v2 = value(r4.z);
                    if (r4.w & 0x80000000u) v3 = value(r4.w);
                    if (r5.x & 0x80000000u) v4 = value(r5.x);
                    if (r5.y & 0x80000000u) v5 = value(r5.y);
                    if (r5.z & 0x80000000u) v6 = value(r5.z);
                    if (r5.w & 0x80000000u) v7 = value(r5.w);
                }
                asm volatile("st.global.L2::evict_first.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(out + k), "r"(v0), "r"(v1), "r"(v2), "r"(v3),
                             "r"(v4), "r"(v5), "r"(v6), "r"(v7)
                             : "memory");
            }
        } else if ((stride & 3u) == 0) {
            for (uint32_t k = 0; k < stride; k += 4) {
                const uint4 r4 = *reinterpret_cast<const uint4*>(res + k);  // entries beyond the outcome's own are "no writer"
                *reinterpret_cast<int4*>(out + k) = make_int4(value(r4.x), value(r4.y), value(r4.z), value(r4.w));
            }
        } else {
            for (uint32_t k = 0; k < stride; ++k) out[k] = value(res[k]);
        }
    };

    // thread 0 takes the ticket of the NEXT tile while the current one is processed (hides the atomic's latency)
    long long ticket_ahead = threadIdx.x == 0 ? static_cast<long long>(atomicAdd(P.ticket, 1u)) : 0;
    long long dbg_t0 = 0, dbg_tiles = 0;
    if (P.debug && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_t0));
    // All CTAs start together, and a tile is a memory-bound burst (prefetch + pre-scan) followed by a long compute
    // phase (the walk): CTAs that stay in lock-step alternate between fighting for HBM and leaving it idle (measured:
    // a stable mode that is 1.4x slower). Spreading the phases once, at the start (each CTA delays its first tile by a
    // pseudo-random fraction of a tile time) removed that mode when the look-back was narrow; with the 128-wide probes
    // and rows that stay pending it does the opposite: the early CTAs block on the look-back of predecessors that are
    // still sleeping and are released together. Measured per launch (100 M lines): with the delay 12.0 ms or 17.6 ms at
    // random, without it 11.85 ms every time (profiles/README.md, round 1e). Kept as a diagnostic switch.
    if ((P.per & 2u) && P.n_tiles >= 8ll * gridDim.x) {  // off by default (GORP_CW_PREFETCH=2 turns it on), see above
        if (threadIdx.x == 0) {
            const unsigned long long delay_ns = static_cast<unsigned long long>((blockIdx.x * 2654435761u) >> 27) * 2500ull;  // 0 .. 77.5 us
            unsigned long long t_start, t_now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
            do {
                __nanosleep(2000);
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_now));
            } while (t_now - t_start < delay_ns);
        }
        __syncthreads();
    }
    if (kDbg) dc_t = clock64();
    for (;;) {
        __syncthreads();  // the previous tile no longer uses s_tile / s_warp; table setup done (first iteration)
        if (threadIdx.x == 0) s_tile = ticket_ahead;
        __syncthreads();
        if (kDbg) dtick(dc_endwait);
        const int64_t tile = s_tile;
        if (tile >= P.n_tiles) break;
        ++dbg_tiles;
        if (threadIdx.x == 0) ticket_ahead = static_cast<long long>(atomicAdd(P.ticket, 1u));
        // ---- newline pre-scan, per warp with coalesced loads: iteration i reads the 32 x 16 bytes of lane i's chunk;
        // a '\n' at position p counts when it starts a line (p + 1 < n_units). Lane i keeps the count and the first.
        const int64_t tile0 = tile * kT * static_cast<int64_t>(C);
        if (!(P.per & 1u)) {  // start fetching this tile's chunks now; the pre-scan consumes them in order (GORP_CW_PREFETCH=1: off)
            const int64_t n0 = tile0 + static_cast<int64_t>(threadIdx.x) * C;
            if (n0 + C <= P.n_units) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(P.text + n0), "n"(C * 2) : "memory");
        }
        uint32_t cnt = 0, first = 0;
        {
            const int64_t wbase = tile0 + static_cast<int64_t>(warp) * 32 * C;
            const bool inside = wbase + 32 * static_cast<int64_t>(C) + 1 <= P.n_units;
            if (wbase < P.n_units) {
#pragma unroll 4
                for (uint32_t i = 0; i < 32; ++i) {
                    const int64_t p = wbase + i * C + lane * 8;
                    uint32_t m;
                    if (inside) {
                        const uint4 v = __ldg(reinterpret_cast<const uint4*>(P.text + p));
                        m = nl_bits4(v.x, v.y) | (nl_bits4(v.z, v.w) << 4);
                    } else {
                        const uint4 v = load_chunk(P.text, p, P.n_units);
                        m = nl_bits4(v.x, v.y) | (nl_bits4(v.z, v.w) << 4);
                        const int64_t room = P.n_units - 1 - p;  // positions (relative to p) that count: [0, room)
                        m = room <= 0 ? 0u : (room < 8 ? m & ((1u << static_cast<uint32_t>(room)) - 1u) : m);
                    }
                    const uint32_t tot = __reduce_add_sync(0xffffffffu, static_cast<uint32_t>(__popc(m)));
                    const uint32_t has = __ballot_sync(0xffffffffu, m != 0);
                    const uint32_t fl = has ? static_cast<uint32_t>(__ffs(has)) - 1u : 0u;
                    const uint32_t mf = __shfl_sync(0xffffffffu, m, fl);
                    if (lane == i) {
                        cnt = tot;
                        first = fl * 8 + static_cast<uint32_t>(__ffs(mf)) - 1u;
                    }
                }
            }
        }
        if (kDbg) dtick(dc_pre);
        const bool line0 = tile == 0 && threadIdx.x == 0 && P.n_units > 0;  // the line at offset 0
        const uint32_t mine = cnt + (line0 ? 1u : 0u);

        // ---- block scan of the line counts; warp 0 then resolves the tile's first row by decoupled look-back
        uint32_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint32_t wbase = 0, total = 0;
        for (uint32_t w = 0; w < n_warps; ++w) {
            const uint32_t x = s_warp[w];
            if (w < warp) wbase += x;
            total += x;
        }
        row0 = wbase + incl - mine;
        base_known = false;
        if (kDbg) dtick(dc_scan);
        if (warp == 0) {
            unsigned long long pre = 0;
            if (tile > 0) {
                if (lane == 0) st_release(P.tile_status + tile, kStAgg | static_cast<unsigned long long>(total));
                // 128 predecessors per probe (4 independent acquire loads per lane): lane l looks at the tiles at distance
                // 4l+1 .. 4l+4. All resident CTAs publish their aggregate at about the same time, so the nearest
                // inclusive prefix can be a whole generation (grid size) away; wide probes keep that to 2-3 round trips.
                for (int64_t j = tile - 1;; j -= 128) {
                    unsigned long long v[4];
#pragma unroll
                    for (int m = 0; m < 4; ++m) {
                        const int64_t idx = j - (lane * 4 + m);
                        v[m] = idx >= 0 ? ld_acquire(P.tile_status + idx) : kStPre;  // before tile 0: an empty inclusive prefix
                    }
#pragma unroll
                    for (int m = 0; m < 4; ++m) {
                        const int64_t idx = j - (lane * 4 + m);
                        while ((v[m] >> 62) == 0) {
                            __nanosleep(32);
                            v[m] = ld_acquire(P.tile_status + idx);
                        }
                    }
                    unsigned long long part = 0;  // aggregates up to and including the lane's nearest inclusive prefix
                    bool has = false;
#pragma unroll
                    for (int m = 0; m < 4; ++m)
                        if (!has) {
                            part += v[m] & ~(3ull << 62);
                            has = (v[m] >> 62) == 2;
                        }
                    const uint32_t pmask = __ballot_sync(0xffffffffu, has);
                    const uint32_t firstp = pmask ? static_cast<uint32_t>(__ffs(pmask)) - 1u : 32u;
                    unsigned long long l = lane <= firstp ? part : 0ull;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
                    pre += l;
                    if (pmask) break;
                }
            }
            if (lane == 0) {
                const long long line_end = static_cast<long long>(pre) + total;
                st_release(P.tile_status + tile, kStPre | static_cast<unsigned long long>(line_end));
                const bool over = line_end > P.cap_lines;
                if (over) atomicOr(reinterpret_cast<unsigned long long*>(P.totals + 2), 1ull);
                if (tile == P.n_tiles - 1) {
                    P.totals[0] = line_end;
                    // line i spans [line_off[i], line_off[i+1] - 1): a text that does not end in '\n' gets n_units + 1
                    if (!over) P.line_off[line_end] = P.n_units + (P.text[P.n_units - 1] == 0x0A ? 0 : 1);
                }
                s_base = static_cast<long long>(pre);
                s_skip_writes = over ? 1 : 0;
                __threadfence_block();
                *reinterpret_cast<volatile unsigned int*>(&s_flag) = static_cast<unsigned int>(tile + 1);
            }
            __syncwarp();
            if (kDbg) dtick(dc_look);
        }

        // ---- walk the owned lines one after the other; positions are tile-relative (pos = unit - tile0)
        for (uint32_t k = 0; k < 2 * A.n_slots; ++k) sts32(slot_abs0 + k * slot_stride, 0x80000000u);  // "never written"
        if (mine) {
            const uint32_t c_end = (threadIdx.x + 1) * C;  // tile-relative end of the chunk
            const int64_t n_rel = P.n_units - tile0;       // tile-relative end of the text
            uint32_t start = line0 ? 0u : threadIdx.x * C + first + 1;
            uint32_t bank_abs = slot_abs0;
            Pending pd{0, 0, 0, 0};
            bool pending = false;
            uint32_t idx = 0, it = 0;
            uint32_t q = start & ~15u;
            uint32_t lo = start & 15u;
            uint32_t ent = lo ? ((A.skip_base + lo - 1) * row_q) << 18 : 0u;
            bool active = true;
            while (active) {
                const Units16 u = load_units16(P.text, tile0 + q, P.n_units);
                if (((u.a.x | u.a.y | u.a.z | u.a.w | u.b.x | u.b.y | u.b.z | u.b.w) & 0xFF80FF80u) == 0u) {
                    one_step<0>(ent, u.a.x, rows_abs, bank_abs, q);
                    one_step<2>(ent, u.a.x, rows_abs, bank_abs, q + 1);
                    one_step<0>(ent, u.a.y, rows_abs, bank_abs, q + 2);
                    one_step<2>(ent, u.a.y, rows_abs, bank_abs, q + 3);
                    one_step<0>(ent, u.a.z, rows_abs, bank_abs, q + 4);
                    one_step<2>(ent, u.a.z, rows_abs, bank_abs, q + 5);
                    one_step<0>(ent, u.a.w, rows_abs, bank_abs, q + 6);
                    one_step<2>(ent, u.a.w, rows_abs, bank_abs, q + 7);
                    one_step<0>(ent, u.b.x, rows_abs, bank_abs, q + 8);
                    one_step<2>(ent, u.b.x, rows_abs, bank_abs, q + 9);
                    one_step<0>(ent, u.b.y, rows_abs, bank_abs, q + 10);
                    one_step<2>(ent, u.b.y, rows_abs, bank_abs, q + 11);
                    one_step<0>(ent, u.b.z, rows_abs, bank_abs, q + 12);
                    one_step<2>(ent, u.b.z, rows_abs, bank_abs, q + 13);
                    one_step<0>(ent, u.b.w, rows_abs, bank_abs, q + 14);
                    one_step<2>(ent, u.b.w, rows_abs, bank_abs, q + 15);
                } else {
                    ent = slow_chunk(A, ent, u.a, P.text, tile0 + q, P.n_units, rows_abs, bank_abs, q, fin_ent);
                    ent = slow_chunk(A, ent, u.b, P.text, tile0 + q + 8, P.n_units, rows_abs, bank_abs, q + 8, fin_ent);
                }
                q += 16;
                if (ent >= fin_ent) {  // the line ended inside these 16 units
                    if (pending) { flush(pd, tile, tile0); if (kDbg) dtick(dc_flush); }  // rare: two line ends within 2 iterations
                    pd.outcome = (((ent >> 18) - A.fin_base * row_q) * inv_row_q) >> 16;  // exact: a multiple of row_q below 2^14
                    pd.idx = idx++;
                    pd.bank_abs = bank_abs;
                    pd.start = start;
                    pending = true;
                    const uint32_t nl = lds32(bank_abs + len_off);  // tile-relative position of the terminating '\n'
                    if (nl < c_end && static_cast<int64_t>(nl) + 1 < n_rel) {
                        start = nl + 1;
                        q = start & ~15u;
                        lo = start & 15u;
                        ent = lo ? ((A.skip_base + lo - 1) * row_q) << 18 : 0u;
                        bank_abs = bank_abs == slot_abs0 ? slot_abs0 + bank_bytes : slot_abs0;
                    } else {
                        active = false;
                    }
                }
                // rows go out every other iteration, once the tile's first row is known (no waiting here: the row can stay
                // pending until the thread's next line ends)
                if ((++it & 1u) == 0 && pending &&
                    (base_known || *reinterpret_cast<volatile unsigned int*>(&s_flag) == static_cast<unsigned int>(tile + 1))) {
                    flush(pd, tile, tile0);
                    if (kDbg) dtick(dc_flush);
                    pending = false;
                }
            }
            if (pending) {
                flush(pd, tile, tile0);
                if (kDbg) dtick(dc_flush);
            }
        }
        if (kDbg) dtick(dc_walk);
    }
    if (kDbg) {
        unsigned long long v[7] = {dc_pre, dc_scan, dc_walk, dc_poll, dc_flush, dc_endwait, dc_look};
        for (int i = 0; i < 7; ++i) {
            unsigned long long x = v[i];
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            if (lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(P.debug) + blockIdx.x * 16 + 4 + i, x);
        }
    }
    if (P.debug && threadIdx.x == 0) {  // GORP_ONEPASS_DEBUG: where and how long this CTA ran
        long long t1;
        uint32_t smid;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        P.debug[blockIdx.x * 16 + 0] = smid;
        P.debug[blockIdx.x * 16 + 1] = dbg_tiles;
        P.debug[blockIdx.x * 16 + 2] = dbg_t0;
        P.debug[blockIdx.x * 16 + 3] = t1;
    }
    if (smem_hist) {
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < n_bins; i += kT)
            if (s_hist[i]) atomicAdd(P.hist + i, static_cast<unsigned long long>(s_hist[i]));
    }
}

}  // namespace

using CwKernel = void (*)(OnePassParams);
static CwKernel cw_kernel(uint32_t, bool dbg) { return dbg ? chunkwalk_kernel<true> : chunkwalk_kernel<false>; }

size_t chunkwalk_smem_bytes(const OnePassDev& a, uint32_t threads) {
    size_t b = static_cast<size_t>(a.n_rows) * a.width * 4;
    b += static_cast<size_t>((a.n_outcomes * a.max_slots + 3) & ~3u) * 4;
    b += 2 * static_cast<size_t>((a.n_outcomes + 3) & ~3u) * 4;
    b += 2 * static_cast<size_t>(a.n_slots) * threads * 4;
    return b + 128;
}

bool k0_chunkwalk_plan(const OnePassDev& a, uint32_t* threads) {
    if (!a.enabled) return false;
    // the CTA size that keeps the most warps resident per SM (ties: the larger CTA, fewer table copies)
    uint32_t best = 0, best_warps = 0;
    uint32_t forced = 0;
    if (const char* f = std::getenv("GORP_CW_THREADS")) forced = static_cast<uint32_t>(std::atoi(f));
    for (uint32_t kT : {512u, 384u, 256u, 128u}) {
        if (forced && kT != forced) continue;
        if (static_cast<uint64_t>(a.n_slots) * kT > 0x3FFFu) continue;  // slot offsets are 14 bits of the table entry
        const size_t smem = chunkwalk_smem_bytes(a, kT);
        if (smem > 226 * 1024) continue;
        allow_max_dynamic_smem(cw_kernel(kT, false));
        allow_max_dynamic_smem(cw_kernel(kT, true));
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cw_kernel(kT, false), static_cast<int>(kT), smem) != cudaSuccess) {
            cudaGetLastError();
            continue;
        }
        const uint32_t warps = static_cast<uint32_t>(per_sm) * kT / 32;
        if (warps > best_warps) best = kT, best_warps = warps;
    }
    if (!best) return false;
    *threads = best;
    return true;
}

int k0_chunkwalk_grid(const Launch& L, const OnePassParams& P, uint32_t threads) {
    const size_t smem = chunkwalk_smem_bytes(P.a, threads);
    allow_max_dynamic_smem(cw_kernel(threads, false));
    allow_max_dynamic_smem(cw_kernel(threads, true));
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cw_kernel(threads, false), static_cast<int>(threads), smem);
    if (per_sm < 1) per_sm = 1;
    const int64_t cap = static_cast<int64_t>(L.sm_count) * per_sm;
    int g = static_cast<int>(P.n_tiles < cap ? P.n_tiles : cap);
    return g < 1 ? 1 : g;
}

void k0_chunkwalk_extract(const Launch& L, const OnePassParams& P, uint32_t threads) {
    const size_t smem = chunkwalk_smem_bytes(P.a, threads);
    const int g = k0_chunkwalk_grid(L, P, threads);
    cw_kernel(threads, P.debug != nullptr)<<<g, static_cast<int>(threads), smem, L.stream>>>(P);
}

}  // namespace gorp
