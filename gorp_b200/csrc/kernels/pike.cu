// Simulating Pike VM (sm_100a): the capture half for extractions whose java.util.regex could not be determinised within
// the limits of host/capture.hpp (state / register / command-list explosion). The reference accepts any regex that
// Pattern.compile accepts (jdkre/JDKRegexpExtractionCooker.java:20-26, JDKRegexpCookedExtraction.java:36-59), so such an
// extraction is not refused: on the table-driven paths it carries a placeholder automaton that accepts nothing — every line
// the combined DFA assigns to it comes out as CAPTURE_FAIL(e) — and this pass then decides exactly those lines by running
// the priority-ordered Pike program itself: thread list = live program counters in java.util.regex preference order, each
// with its capture slots; acceptance only at the end of the line; the first thread that reaches MATCH there wins.
// One thread per line, thread lists in global scratch memory: slow by construction (O(line x program)), and rare.
#include "device_common.cuh"

namespace gorp {

namespace {

constexpr int kPikeThreads = 128;

__global__ void __launch_bounds__(kPikeThreads) pike_fixup_kernel(PikeDev P, const uint16_t* __restrict__ text, const int64_t* __restrict__ line_off,
                                                                 int sep, int64_t n_lines, uint32_t span_stride, int32_t* __restrict__ ext_id,
                                                                 int32_t* __restrict__ spans, unsigned long long* __restrict__ hist, uint32_t n_ext,
                                                                 int32_t* __restrict__ scratch) {
    const uint32_t ent = 1u + P.max_slots;  // ints per thread-list entry: pc, capture slots
    const size_t per_thread = static_cast<size_t>(2) * P.max_insts * ent + P.max_insts;
    int32_t* mine = scratch + (static_cast<size_t>(blockIdx.x) * kPikeThreads + threadIdx.x) * per_thread;
    int32_t* list[2] = {mine, mine + static_cast<size_t>(P.max_insts) * ent};
    uint32_t* seen = reinterpret_cast<uint32_t*>(mine + static_cast<size_t>(2) * P.max_insts * ent);
    uint32_t gen = 0;
    bool seen_clean = false;
    for (int64_t line = static_cast<int64_t>(blockIdx.x) * kPikeThreads + threadIdx.x; line < n_lines; line += static_cast<int64_t>(gridDim.x) * kPikeThreads) {
        const int32_t code = ext_id[line];
        if (code > -2) continue;
        const uint32_t e = static_cast<uint32_t>(-2 - code);
        const PikeExtDev X = P.ext[e];
        if (!X.enabled) continue;
        if (!seen_clean) {
            for (uint32_t i = 0; i < P.max_insts; ++i) seen[i] = 0;
            seen_clean = true;
        }
        const int64_t a = line_off[line], b = line_off[line + 1] - sep;
        int cur = 0;
        uint32_t n_cur = 1;
        list[0][0] = 0;
        for (uint32_t k = 0; k < X.n_slots; ++k) list[0][1 + k] = -1;
        for (int64_t p = a; p < b && n_cur; ++p) {
            const uint32_t u = text[p];
            uint32_t c = __ldg(P.cls + u);
            if ((u & 0xFC00u) == 0xD800u && p + 1 < b && (text[p + 1] & 0xFC00u) == 0xDC00u) c = P.pair_hi_class;
            ++gen;
            uint32_t n_nxt = 0;
            int32_t* from = list[cur];
            int32_t* to = list[cur ^ 1];
            for (uint32_t i = 0; i < n_cur; ++i) {
                const int32_t pc = from[i * ent];
                for (uint32_t j = P.clo_off[X.inst_off + pc], j1 = P.clo_off[X.inst_off + pc + 1]; j < j1; ++j) {
                    const int32_t t = P.clo_target[j];
                    if (t < 0 || seen[t] == gen) continue;
                    seen[t] = gen;
                    if (!P.accepts[X.acc_off + static_cast<size_t>(t) * P.n_classes + c]) continue;
                    int32_t* dst = to + n_nxt * ent;
                    dst[0] = t + 1;
                    const unsigned long long mask = P.clo_mask[j];
                    for (uint32_t k = 0; k < X.n_slots; ++k) dst[1 + k] = (mask >> k) & 1ull ? static_cast<int32_t>(p - a) : from[i * ent + 1 + k];
                    ++n_nxt;
                }
            }
            cur ^= 1;
            n_cur = n_nxt;
        }
        // end of line: the first thread (in preference order) whose closure reaches MATCH
        bool matched = false;
        const int32_t* from = list[cur];
        for (uint32_t i = 0; i < n_cur && !matched; ++i) {
            const int32_t pc = from[i * ent];
            for (uint32_t j = P.clo_off[X.inst_off + pc], j1 = P.clo_off[X.inst_off + pc + 1]; j < j1; ++j) {
                if (P.clo_target[j] >= 0) continue;
                const unsigned long long mask = P.clo_mask[j];
                int32_t* out = spans + line * span_stride;
                for (uint32_t k = 0; k < span_stride; ++k)
                    out[k] = k < X.n_slots ? ((mask >> k) & 1ull ? static_cast<int32_t>(b - a) : from[i * ent + 1 + k]) : -1;
                matched = true;
                break;
            }
        }
        if (matched) {
            ext_id[line] = static_cast<int32_t>(e);
            atomicAdd(hist + e, 1ull);
            atomicAdd(hist + n_ext + 1, ~0ull);  // it was counted as a capture failure
        }
    }
}

}  // namespace

size_t pike_scratch_ints_per_thread(const PikeDev& p) { return static_cast<size_t>(2) * p.max_insts * (1u + p.max_slots) + p.max_insts; }

void k_pike_fixup(const Launch& L, const PikeDev& P, const uint16_t* text, const int64_t* line_off, int sep, int64_t n_lines, uint32_t span_stride,
                  int32_t* ext_id, int32_t* spans, unsigned long long* hist, uint32_t n_ext, int32_t* scratch, uint32_t scratch_threads) {
    if (n_lines <= 0 || scratch_threads < kPikeThreads) return;
    const int64_t want = (n_lines + kPikeThreads - 1) / kPikeThreads, cap = scratch_threads / kPikeThreads;
    pike_fixup_kernel<<<static_cast<int>(want < cap ? want : cap), kPikeThreads, 0, L.stream>>>(P, text, line_off, sep, n_lines, span_stride, ext_id, spans,
                                                                                             hist, n_ext, scratch);
}

}  // namespace gorp
