// BENCH / TEST INFRASTRUCTURE: libgorpgen.so — host loops and device kernels around csrc/tools/corpusgen.h (one source for
// both, see there). Not linked into libgorpcuda.so; the product never calls it.
#include <cuda_runtime.h>

#include <cstring>
#include <thread>
#include <vector>

#include "corpusgen.h"

struct HostProgram {  // the C ABI's view (all host pointers)
    const uint32_t* kinds;
    const int32_t* ops;
    const uint32_t* choice;
    const uint16_t* strings;
    const uint16_t* qtable;
    uint32_t n_kinds, n_ops, n_choice, n_strings;
    uint32_t params[16];  // p_outlier, p_diverge, p_suppl, p_nonascii, outlier_len, diverge_off, diverge_n, suppl_off, suppl_n,
                          // nonascii_off, nonascii_n, alnum_off, alnum_n, outlier_alpha_off, outlier_alpha_n, reserved
};

namespace {

CgProgram view(const HostProgram& h, const uint32_t* kinds, const int32_t* ops, const uint32_t* choice, const uint16_t* strings,
               const uint16_t* qtable) {
    CgProgram P{};
    P.kinds = kinds;
    P.ops = ops;
    P.choice = choice;
    P.strings = strings;
    P.qtable = qtable;
    P.n_kinds = h.n_kinds;
    P.p_outlier = h.params[0];
    P.p_diverge = h.params[1];
    P.p_suppl = h.params[2];
    P.p_nonascii = h.params[3];
    P.outlier_len = h.params[4];
    P.diverge_off = h.params[5];
    P.diverge_n = h.params[6];
    P.suppl_off = h.params[7];
    P.suppl_n = h.params[8];
    P.nonascii_off = h.params[9];
    P.nonascii_n = h.params[10];
    P.alnum_off = h.params[11];
    P.alnum_n = h.params[12];
    P.outlier_alpha_off = h.params[13];
    P.outlier_alpha_n = h.params[14];
    return P;
}

__global__ void lengths_kernel(CgProgram P, uint64_t seed, int64_t first, int64_t n, int32_t* len) {
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
        len[i] = static_cast<int32_t>(cg_line(P, seed, static_cast<uint64_t>(first + i), nullptr));
}

__global__ void fill_kernel(CgProgram P, uint64_t seed, int64_t first, int64_t n, const int64_t* off, uint16_t* out) {
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t at = off[i];
        const int64_t len = cg_line(P, seed, static_cast<uint64_t>(first + i), out + at);
        out[at + len] = 0x0A;
    }
}

struct DeviceProgram {
    void* p[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    CgProgram P{};
    bool ok = false;
    DeviceProgram(const HostProgram& h, cudaStream_t s) {
        const void* src[5] = {h.kinds, h.ops, h.choice, h.strings, h.qtable};
        const size_t bytes[5] = {h.n_kinds * 8ull, h.n_ops * 16ull, h.n_choice * 12ull, h.n_strings * 2ull, 512};
        for (int i = 0; i < 5; ++i) {
            if (cudaMalloc(&p[i], bytes[i] + 16) != cudaSuccess) return;
            if (cudaMemcpyAsync(p[i], src[i], bytes[i], cudaMemcpyHostToDevice, s) != cudaSuccess) return;
        }
        P = view(h, static_cast<const uint32_t*>(p[0]), static_cast<const int32_t*>(p[1]), static_cast<const uint32_t*>(p[2]),
                 static_cast<const uint16_t*>(p[3]), static_cast<const uint16_t*>(p[4]));
        ok = true;
    }
    ~DeviceProgram() {
        for (void* q : p)
            if (q) cudaFree(q);
    }
};

template <class F>
void parallel_for(int64_t n, int threads, F&& f) {
    if (threads < 1) threads = 1;
    if (threads == 1 || n < 4096) {
        f(0, n);
        return;
    }
    std::vector<std::thread> th;
    for (int t = 0; t < threads; ++t) th.emplace_back([&, t]() { f(n * t / threads, n * (t + 1) / threads); });
    for (auto& x : th) x.join();
}

}  // namespace

extern "C" {

int cg_lengths(const HostProgram* h, uint64_t seed, int64_t first, int64_t n, int32_t* len, int threads) {
    const CgProgram P = view(*h, h->kinds, h->ops, h->choice, h->strings, h->qtable);
    parallel_for(n, threads, [&](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; ++i) len[i] = static_cast<int32_t>(cg_line(P, seed, static_cast<uint64_t>(first + i), nullptr));
    });
    return 0;
}

// off[i] = unit offset of line i in `out` (off[i + 1] - off[i] = length + 1: every line is followed by '\n')
int cg_fill(const HostProgram* h, uint64_t seed, int64_t first, int64_t n, const int64_t* off, uint16_t* out, int threads) {
    const CgProgram P = view(*h, h->kinds, h->ops, h->choice, h->strings, h->qtable);
    parallel_for(n, threads, [&](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; ++i) {
            const int64_t len = cg_line(P, seed, static_cast<uint64_t>(first + i), out + off[i]);
            out[off[i] + len] = 0x0A;
        }
    });
    return 0;
}

int cg_lengths_device(const HostProgram* h, uint64_t seed, int64_t first, int64_t n, int32_t* d_len, void* stream) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    DeviceProgram D(*h, s);
    if (!D.ok) return -1;
    if (n > 0) lengths_kernel<<<static_cast<int>(n / 256 + 1 < 65535 * 16 ? n / 256 + 1 : 65535 * 16), 256, 0, s>>>(D.P, seed, first, n, d_len);
    return cudaStreamSynchronize(s) == cudaSuccess && cudaGetLastError() == cudaSuccess ? 0 : -2;
}

int cg_fill_device(const HostProgram* h, uint64_t seed, int64_t first, int64_t n, const int64_t* d_off, uint16_t* d_out, void* stream) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    DeviceProgram D(*h, s);
    if (!D.ok) return -1;
    if (n > 0) fill_kernel<<<static_cast<int>(n / 256 + 1 < 65535 * 16 ? n / 256 + 1 : 65535 * 16), 256, 0, s>>>(D.P, seed, first, n, d_off, d_out);
    return cudaStreamSynchronize(s) == cudaSuccess && cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // extern "C"
