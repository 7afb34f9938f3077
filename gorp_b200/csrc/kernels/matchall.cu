// matchAll (sm_100a): PolyMatcher.match for a batch of strings with the FULL accept lists
// (reference autom/PolyMatcher.java:123-133 -> Automata.accept, autom/Automata.java:137-139: every regex index that
// accepts the whole string, ascending) — the secondary entry point Gorp.getMatcher().match(s) (Gorp.java:135-137) that
// MultiPatternTest pins. It walks the reference's own tables (Automata._alphabet / _transitions / _accept as the blob
// carries them), one line per thread; the extract path never needs more than the first index, so this one is not tuned.
//   pass 1  final state per line (or -1), number of accepting indexes per line
//   scan    exclusive prefix -> accept_off (CSR)
//   pass 2  copies the accept lists
#include "device_common.cuh"

namespace gorp {

namespace {

constexpr int kMaThreads = 256;

__global__ void __launch_bounds__(kMaThreads) matchall_walk_kernel(MatchAllDev d, const uint16_t* __restrict__ text, const int64_t* __restrict__ off,
                                                                   int64_t n_lines, int32_t* __restrict__ state, uint32_t* __restrict__ count) {
    for (int64_t line = static_cast<int64_t>(blockIdx.x) * kMaThreads + threadIdx.x; line < n_lines; line += static_cast<int64_t>(gridDim.x) * kMaThreads) {
        const int64_t a = off[line], b = off[line + 1];
        int32_t p = 0;
        for (int64_t i = a; i < b && p >= 0; ++i) p = __ldg(d.trans + static_cast<int64_t>(p) * d.n_classes + __ldg(d.classmap + __ldg(text + i)));
        state[line] = p;
        count[line] = p < 0 ? 0u : __ldg(d.accept_off + p + 1) - __ldg(d.accept_off + p);
    }
}

__global__ void __launch_bounds__(kMaThreads) matchall_fill_kernel(MatchAllDev d, const int32_t* __restrict__ state, const int64_t* __restrict__ out_off,
                                                                   int64_t n_lines, int32_t* __restrict__ out) {
    for (int64_t line = static_cast<int64_t>(blockIdx.x) * kMaThreads + threadIdx.x; line < n_lines; line += static_cast<int64_t>(gridDim.x) * kMaThreads) {
        const int32_t p = state[line];
        if (p < 0) continue;
        const uint32_t a0 = __ldg(d.accept_off + p), a1 = __ldg(d.accept_off + p + 1);
        int32_t* dst = out + out_off[line];
        for (uint32_t i = a0; i < a1; ++i) dst[i - a0] = __ldg(d.accept_list + i);
    }
}

}  // namespace

void k_matchall_walk(const Launch& L, const MatchAllDev& d, const uint16_t* text, const int64_t* off, int64_t n_lines, int32_t* state, uint32_t* count) {
    if (n_lines <= 0) return;
    const int64_t want = (n_lines + kMaThreads - 1) / kMaThreads, cap = static_cast<int64_t>(L.sm_count) * 8;
    matchall_walk_kernel<<<static_cast<int>(want < cap ? want : cap), kMaThreads, 0, L.stream>>>(d, text, off, n_lines, state, count);
}

void k_matchall_fill(const Launch& L, const MatchAllDev& d, const int32_t* state, const int64_t* out_off, int64_t n_lines, int32_t* out) {
    if (n_lines <= 0) return;
    const int64_t want = (n_lines + kMaThreads - 1) / kMaThreads, cap = static_cast<int64_t>(L.sm_count) * 8;
    matchall_fill_kernel<<<static_cast<int>(want < cap ? want : cap), kMaThreads, 0, L.stream>>>(d, state, out_off, n_lines, out);
}

}  // namespace gorp
