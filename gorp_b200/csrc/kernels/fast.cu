// Fast tiers of the hot kernels (sm_100a): no per-unit masking, no class lookup for ASCII chunks.
#include "device_common.cuh"

namespace gorp {
namespace {

constexpr int kFastThreads = 512;

using namespace dev;

__global__ void __launch_bounds__(kFastThreads) dfa_direct_kernel(DfaDirectDev d, const uint16_t* __restrict__ text,
                                                                  const int64_t* __restrict__ line_off, int64_t n_lines,
                                                                  int32_t* __restrict__ ext_id) {
    extern __shared__ __align__(16) uint32_t s_rows[];
    const uint32_t tbase = static_cast<uint32_t>(__cvta_generic_to_shared(s_rows));
    for (uint32_t i = threadIdx.x; i < d.n_rows * 128u; i += kFastThreads) s_rows[i] = tbase + (__ldg(d.rows + i) << 9);
    __syncthreads();
    const uint32_t fin_abs = tbase + (d.fin_base << 9);
    const uint32_t skip_abs = tbase + (d.skip_base << 9);
    for (int64_t line = static_cast<int64_t>(blockIdx.x) * kFastThreads + threadIdx.x; line < n_lines;
         line += static_cast<int64_t>(gridDim.x) * kFastThreads) {
        const int64_t a = line_off[line];
        const uint16_t* p = text + (a & ~int64_t(7));
        const uint32_t lo = static_cast<uint32_t>(a) & 7u;
        uint32_t st = lo ? skip_abs + ((lo - 1) << 9) : tbase;
        do {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
            p += 8;
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
                st = dfa_slow_chunk(d, tbase, st, v);
            }
        } while (st < fin_abs);
        const int32_t e = static_cast<int32_t>((st - fin_abs) >> 9) - 1;
        ext_id[line] = e;
    }
}

// ------------------------------------------------------------------ K4 fast tier
constexpr int kCapThreads = kCapFastThreads;

__global__ void __launch_bounds__(kCapThreads, 2) tdfa_fast_kernel(TdfaFastDev f, CapDev c, const uint16_t* __restrict__ text,
                                                                   int64_t n_units,
                                                                   const int64_t* __restrict__ line_off, int64_t n_lines,
                                                                uint32_t span_stride,
                                                                int32_t* __restrict__ ext_id, int32_t* __restrict__ spans) {
    extern __shared__ __align__(16) uint32_t s_img[];
    for (uint32_t i = threadIdx.x; i < f.image_words; i += kCapThreads) s_img[i] = __ldg(f.image + i);
    __syncthreads();
    const uint32_t cls_abs = static_cast<uint32_t>(__cvta_generic_to_shared(s_img));
    const uint32_t reg_stride = kCapThreads * 4;
    const uint32_t reg_abs = cls_abs + f.image_words * 4 + threadIdx.x * 4;
    const uint32_t len_off = (f.n_regs + 1) * reg_stride;
    for (int64_t line = static_cast<int64_t>(blockIdx.x) * kCapThreads + threadIdx.x; line < n_lines;
         line += static_cast<int64_t>(gridDim.x) * kCapThreads) {
        const int32_t e = ext_id[line];
        int32_t* out = spans + line * span_stride;
        if (e < 0) {
            for (uint32_t k = 0; k < span_stride; ++k) out[k] = -1;
            continue;
        }
        const FastExtDev fx = f.ext[e];
        const ExtDev x = c.ext[e];
        const uint32_t tab_abs = cls_abs + fx.tab_off;
        const int64_t a = line_off[line];
        int64_t q = a & ~int64_t(7);
        const uint32_t lo = static_cast<uint32_t>(a - q);
        uint32_t st = lo ? (fx.n_states + lo - 1) * fx.row_bytes : 0u;
        uint32_t pos = static_cast<uint32_t>(q - a);  // negative while skipping: only ever stored to the dummy register
        for (;;) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(text + q));
            const uint32_t st0 = st;
            if (((v.x | v.y | v.z | v.w) & 0xFF80FF80u) == 0u) {
                cap_step<0>(st, v.x, cls_abs, tab_abs, reg_abs, pos);
                cap_step<2>(st, v.x, cls_abs, tab_abs, reg_abs, pos + 1);
                cap_step<0>(st, v.y, cls_abs, tab_abs, reg_abs, pos + 2);
                cap_step<2>(st, v.y, cls_abs, tab_abs, reg_abs, pos + 3);
                cap_step<0>(st, v.z, cls_abs, tab_abs, reg_abs, pos + 4);
                cap_step<2>(st, v.z, cls_abs, tab_abs, reg_abs, pos + 5);
                cap_step<0>(st, v.w, cls_abs, tab_abs, reg_abs, pos + 6);
                cap_step<2>(st, v.w, cls_abs, tab_abs, reg_abs, pos + 7);
                if (st == fx.slow_off) st = tdfa_slow_chunk(c, x, fx, st0, text, q, a, n_units, reg_abs, reg_stride, len_off);
            } else {
                st = tdfa_slow_chunk(c, x, fx, st0, text, q, a, n_units, reg_abs, reg_stride, len_off);
            }
            if (st >= fx.dead_off) break;
            q += 8;
            pos += 8;
        }
        bool ok = st >= fx.frz_off;
        uint32_t s = 0;
        if (ok) {
            s = (st - fx.frz_off) / fx.row_bytes;
            ok = __ldg(c.tdfa_accepting + x.acc_off + s) != 0;
        }
        if (!ok) {
            ext_id[line] = -2 - e;
            for (uint32_t k = 0; k < span_stride; ++k) out[k] = -1;
            continue;
        }
        const uint8_t* __restrict__ fin = c.tdfa_fin + x.fin_off + s * x.n_slots;
        const int32_t len = static_cast<int32_t>(lds32(reg_abs + len_off));
        for (uint32_t k = 0; k < x.n_slots; ++k) {
            const uint32_t r = __ldg(fin + k);
            out[k] = r == 0xFFu ? -1 : (r == 0xFEu ? len : static_cast<int32_t>(lds32(reg_abs + r * reg_stride)));
        }
        for (uint32_t k = x.n_slots; k < span_stride; ++k) out[k] = -1;
    }
}

}  // namespace

void k2_dfa_direct(const Launch& L, const DfaDirectDev& d, const uint16_t* text, const int64_t* line_off, int64_t n_lines,
                   int32_t* ext_id) {
    if (n_lines <= 0) return;
    const size_t smem = static_cast<size_t>(d.n_rows) * 512;
    if (smem > 40 * 1024)
        allow_max_dynamic_smem(dfa_direct_kernel);
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dfa_direct_kernel, kFastThreads, smem);
    if (per_sm < 1) per_sm = 1;
    int64_t want = (n_lines + kFastThreads - 1) / kFastThreads;
    int64_t cap = static_cast<int64_t>(L.sm_count) * per_sm;
    int g = static_cast<int>(want < cap ? want : cap);
    dfa_direct_kernel<<<g, kFastThreads, smem, L.stream>>>(d, text, line_off, n_lines, ext_id);
}

}  // namespace gorp

namespace gorp {
void k4_tdfa_fast(const Launch& L, const TdfaFastDev& f, const CapDev& c, const uint16_t* text, int64_t n_units,
                  const int64_t* line_off, int64_t n_lines, uint32_t span_stride, int32_t* ext_id, int32_t* spans) {
    if (n_lines <= 0) return;
    const size_t smem = static_cast<size_t>(f.image_words) * 4 + static_cast<size_t>(f.n_regs + 2) * kCapThreads * 4;
    if (smem > 40 * 1024)
        allow_max_dynamic_smem(tdfa_fast_kernel);
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tdfa_fast_kernel, kCapThreads, smem);
    if (per_sm < 1) per_sm = 1;
    int64_t want = (n_lines + kCapThreads - 1) / kCapThreads;
    int64_t cap = static_cast<int64_t>(L.sm_count) * per_sm;
    int g = static_cast<int>(want < cap ? want : cap);
    tdfa_fast_kernel<<<g, kCapThreads, smem, L.stream>>>(f, c, text, n_units, line_off, n_lines, span_stride, ext_id, spans);
}
}  // namespace gorp
