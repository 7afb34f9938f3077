// UTF-8 ingest (sm_100a): log files are UTF-8 on disk (the reference reads its definitions that way,
// io/InputLineReader.java:51; callers of Gorp.extract hold the decoded Strings). gorp_extract_text_utf8 ships the BYTES
// over PCIe — one byte per ASCII character instead of two — and decodes them to UTF-16 on the device, which is what
// `new String(bytes, UTF_8)` would have produced on the host; the result offsets are UTF-16 units of that decoded text.
//
//   pass 1  utf8_count   per 4096-byte tile: UTF-16 units its bytes start (0 for a continuation byte, 2 for a 4-byte lead,
//                        else 1) + validation of every sequence (Unicode table 3-7: no overlongs, no surrogates, <= U+10FFFF,
//                        no truncated sequence); the first malformed byte offset is reported (atomicMin)
//           scan         exclusive prefix of the tile counts (scan_u32_to_i64)
//   pass 2  utf8_write   block scan inside the tile, every thread decodes the sequences that START in its 16 bytes
// Both passes are plain HBM streams (1 byte read per unit in pass 1, 1 byte read + 2 bytes written in pass 2).
#include "device_common.cuh"

namespace gorp {

namespace {

constexpr int kUtf8Threads = 256;
constexpr int kUtf8PerThread = 16;
constexpr int kUtf8Tile = kUtf8Threads * kUtf8PerThread;

__device__ __forceinline__ uint32_t byte_at(const uint8_t* __restrict__ src, int64_t i, int64_t n) {
    return i < n ? static_cast<uint32_t>(__ldg(src + i)) : 0x100u;  // 0x100: "no byte" (never a continuation byte)
}

// length of the well-formed sequence that starts with lead byte b0 at position i (1..4), or 0 when malformed
__device__ __forceinline__ uint32_t sequence_length(const uint8_t* __restrict__ src, int64_t i, int64_t n, uint32_t b0) {
    if (b0 < 0x80u) return 1;
    if (b0 < 0xC2u) return 0;  // continuation byte in lead position, or an overlong 2-byte lead (C0, C1)
    const uint32_t b1 = byte_at(src, i + 1, n);
    if (b0 < 0xE0u) return (b1 & 0xC0u) == 0x80u && b1 < 0x100u ? 2u : 0u;
    const uint32_t b2 = byte_at(src, i + 2, n);
    if (b0 < 0xF0u) {
        const uint32_t lo = b0 == 0xE0u ? 0xA0u : 0x80u, hi = b0 == 0xEDu ? 0x9Fu : 0xBFu;  // no overlongs, no surrogates
        return b1 >= lo && b1 <= hi && (b2 & 0xC0u) == 0x80u && b2 < 0x100u ? 3u : 0u;
    }
    if (b0 > 0xF4u) return 0;
    const uint32_t b3 = byte_at(src, i + 3, n);
    const uint32_t lo = b0 == 0xF0u ? 0x90u : 0x80u, hi = b0 == 0xF4u ? 0x8Fu : 0xBFu;  // U+10000 .. U+10FFFF
    return b1 >= lo && b1 <= hi && (b2 & 0xC0u) == 0x80u && b2 < 0x100u && (b3 & 0xC0u) == 0x80u && b3 < 0x100u ? 4u : 0u;
}

__device__ __forceinline__ uint32_t block_sum(uint32_t v, uint32_t* s_warp) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = v;
    __syncthreads();
    uint32_t t = 0;
    for (int w = 0; w < kUtf8Threads / 32; ++w) t += s_warp[w];
    __syncthreads();
    return t;
}

__global__ void __launch_bounds__(kUtf8Threads) utf8_count_kernel(const uint8_t* __restrict__ src, int64_t n, uint32_t* __restrict__ tile_counts,
                                                                 unsigned long long* __restrict__ first_bad) {
    __shared__ uint32_t s_warp[kUtf8Threads / 32];
    const int64_t n_tiles = (n + kUtf8Tile - 1) / kUtf8Tile;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t p0 = tile * kUtf8Tile + static_cast<int64_t>(threadIdx.x) * kUtf8PerThread;
        uint32_t units = 0;
        if (p0 < n) {
            __align__(16) uint8_t b[kUtf8PerThread];
            if (p0 + kUtf8PerThread <= n) {
                *reinterpret_cast<uint4*>(b) = __ldg(reinterpret_cast<const uint4*>(src + p0));
            } else {
                for (int k = 0; k < kUtf8PerThread; ++k) b[k] = p0 + k < n ? __ldg(src + p0 + k) : 0x80;  // padding: counts nothing
            }
            // a continuation byte is legal only inside a sequence: the lead that owns it is checked by the thread that holds
            // the lead, so here every NON-continuation byte is validated as the start of a sequence, and a continuation byte
            // must be covered by the sequence that precedes it (checked by walking the sequences of this thread's range)
            int64_t expect_until = p0;  // continuation bytes before this position belong to a sequence begun earlier
            {   // how far does a sequence begun before p0 reach? look back up to 3 bytes for its lead
                for (int back = 1; back <= 3 && p0 - back >= 0; ++back) {
                    const uint32_t pb = __ldg(src + p0 - back);
                    if ((pb & 0xC0u) != 0x80u) {
                        const uint32_t len = sequence_length(src, p0 - back, n, pb);
                        if (len > static_cast<uint32_t>(back)) expect_until = p0 - back + len;
                        break;
                    }
                }
            }
            for (int k = 0; k < kUtf8PerThread; ++k) {
                const int64_t p = p0 + k;
                if (p >= n) break;
                const uint32_t c = b[k];
                if ((c & 0xC0u) == 0x80u) {
                    if (p >= expect_until) atomicMin(first_bad, static_cast<unsigned long long>(p));  // stray continuation byte
                    continue;
                }
                if (p < expect_until) {  // a lead byte inside an unfinished sequence (the lead's own check reports it)
                    continue;
                }
                const uint32_t len = sequence_length(src, p, n, c);
                if (len == 0) atomicMin(first_bad, static_cast<unsigned long long>(p));
                expect_until = p + (len ? len : 1);
                units += c >= 0xF0u ? 2u : 1u;
            }
        }
        const uint32_t total = block_sum(units, s_warp);
        if (threadIdx.x == 0) tile_counts[tile] = total;
    }
}

__global__ void __launch_bounds__(kUtf8Threads) utf8_write_kernel(const uint8_t* __restrict__ src, int64_t n, const int64_t* __restrict__ tile_base,
                                                                 uint16_t* __restrict__ dst) {
    __shared__ uint32_t s_warp[kUtf8Threads / 32];
    const int64_t n_tiles = (n + kUtf8Tile - 1) / kUtf8Tile;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t p0 = tile * kUtf8Tile + static_cast<int64_t>(threadIdx.x) * kUtf8PerThread;
        __align__(16) uint8_t b[kUtf8PerThread];
        uint32_t units = 0;
        if (p0 < n) {
            if (p0 + kUtf8PerThread <= n) {
                *reinterpret_cast<uint4*>(b) = __ldg(reinterpret_cast<const uint4*>(src + p0));
            } else {
                for (int k = 0; k < kUtf8PerThread; ++k) b[k] = p0 + k < n ? __ldg(src + p0 + k) : 0x80;
            }
            for (int k = 0; k < kUtf8PerThread; ++k) units += (b[k] & 0xC0u) == 0x80u ? 0u : (b[k] >= 0xF0u ? 2u : 1u);
        }
        // exclusive prefix of `units` over the threads of the tile
        uint32_t incl = units;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= static_cast<uint32_t>(o)) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint32_t before = 0;
        for (uint32_t w = 0; w < warp; ++w) before += s_warp[w];
        __syncthreads();
        if (p0 < n) {
            int64_t at = tile_base[tile] + before + incl - units;
            for (int k = 0; k < kUtf8PerThread; ++k) {
                const uint32_t c = b[k];
                if ((c & 0xC0u) == 0x80u) continue;
                const int64_t p = p0 + k;
                if (p >= n) break;
                auto next = [&](int j) -> uint32_t {  // continuation byte j of the sequence (validated by pass 1)
                    return (k + j < kUtf8PerThread ? b[k + j] : byte_at(src, p + j, n)) & 0x3Fu;
                };
                if (c < 0x80u) {
                    dst[at++] = static_cast<uint16_t>(c);
                } else if (c < 0xE0u) {
                    dst[at++] = static_cast<uint16_t>(((c & 0x1Fu) << 6) | next(1));
                } else if (c < 0xF0u) {
                    dst[at++] = static_cast<uint16_t>(((c & 0x0Fu) << 12) | (next(1) << 6) | next(2));
                } else {
                    const uint32_t cp = (((c & 0x07u) << 18) | (next(1) << 12) | (next(2) << 6) | next(3)) - 0x10000u;
                    dst[at++] = static_cast<uint16_t>(0xD800u + (cp >> 10));
                    dst[at++] = static_cast<uint16_t>(0xDC00u + (cp & 0x3FFu));
                }
            }
        }
    }
}

}  // namespace

int64_t utf8_tiles(int64_t n_bytes) { return (n_bytes + kUtf8Tile - 1) / kUtf8Tile; }

void k_utf8_count(const Launch& L, const uint8_t* src, int64_t n, uint32_t* tile_counts, unsigned long long* first_bad) {
    if (n <= 0) return;
    const int64_t tiles = utf8_tiles(n), cap = static_cast<int64_t>(L.sm_count) * 8;
    utf8_count_kernel<<<static_cast<int>(tiles < cap ? tiles : cap), kUtf8Threads, 0, L.stream>>>(src, n, tile_counts, first_bad);
}

void k_utf8_write(const Launch& L, const uint8_t* src, int64_t n, const int64_t* tile_base, uint16_t* dst) {
    if (n <= 0) return;
    const int64_t tiles = utf8_tiles(n), cap = static_cast<int64_t>(L.sm_count) * 8;
    utf8_write_kernel<<<static_cast<int>(tiles < cap ? tiles : cap), kUtf8Threads, 0, L.stream>>>(src, n, tile_base, dst);
}

}  // namespace gorp
