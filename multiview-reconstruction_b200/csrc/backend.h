// Device backend used by the engine.
//
//  * default            : CUDA runtime (the product; sm_100a)
//  * -DMVD_HOST_EMU     : TEST-ONLY host emulation.  The very same kernel bodies (fft_passes.cuh) are executed
//                         thread-by-thread on the CPU so that index math, tiling and the pass schedule can be unit
//                         tested in a container without a GPU.  It is built into tests/host/libmvdecon_hostemu.so,
//                         never into the product library, and the python package cannot load it.
#pragma once
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "fft_passes.cuh"

#ifndef MVD_HOST_EMU
#include <cuda_runtime.h>
#endif

namespace mvd {

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

#ifndef MVD_HOST_EMU
typedef cudaStream_t stream_t;
#define MVD_CUDA_CHECK(call)                                                                         \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess)                                                                      \
            throw ::mvd::Error(std::string(#call) + ": " + cudaGetErrorString(e__) + " (" __FILE__ ":" + \
                               std::to_string(__LINE__) + ")");                                      \
    } while (0)

struct DevExec {
    double s_;
    float m_;
    template <class F> __device__ __forceinline__ void phase(F&& f) { f((int)threadIdx.x); __syncthreads(); }
    __device__ __forceinline__ void stash(int, double s, float m) { s_ = s; m_ = m; }
    __device__ __forceinline__ void unstash(int, double& s, float& m) { s = s_; m = m_; }
};

namespace dev {
inline void set_device(int d) { MVD_CUDA_CHECK(cudaSetDevice(d)); }
inline void* alloc(size_t n) { void* p = nullptr; MVD_CUDA_CHECK(cudaMalloc(&p, n ? n : 1)); return p; }
inline void free_(void* p) { if (p) cudaFree(p); }
inline void h2d(void* d, const void* h, size_t n, stream_t s) { MVD_CUDA_CHECK(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s)); }
inline void d2h(void* h, const void* d, size_t n, stream_t s) { MVD_CUDA_CHECK(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s)); }
inline void d2d(void* d, const void* s_, size_t n, stream_t s) { MVD_CUDA_CHECK(cudaMemcpyAsync(d, s_, n, cudaMemcpyDeviceToDevice, s)); }
inline void zero(void* d, size_t n, stream_t s) { MVD_CUDA_CHECK(cudaMemsetAsync(d, 0, n, s)); }
inline void sync(stream_t s) { MVD_CUDA_CHECK(cudaStreamSynchronize(s)); }
inline stream_t stream_create() { cudaStream_t s; MVD_CUDA_CHECK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking)); return s; }
inline stream_t stream_create_high_priority() {
    int lo = 0, hi = 0;
    MVD_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    cudaStream_t s;
    MVD_CUDA_CHECK(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, hi));
    return s;
}
// clears `height` runs of `width` bytes that start `pitch` bytes apart
inline void zero2d(void* d, size_t pitch, size_t width, size_t height, stream_t s) {
    if (width && height) MVD_CUDA_CHECK(cudaMemset2DAsync(d, pitch, 0, width, height, s));
}
inline void stream_destroy(stream_t s) { if (s) cudaStreamDestroy(s); }
typedef cudaEvent_t event_t;
inline event_t event_create() { cudaEvent_t e; MVD_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); return e; }
inline void event_destroy(event_t e) { if (e) cudaEventDestroy(e); }
inline void event_record(event_t e, stream_t s) { MVD_CUDA_CHECK(cudaEventRecord(e, s)); }
inline void stream_wait(stream_t s, event_t e) { MVD_CUDA_CHECK(cudaStreamWaitEvent(s, e, 0)); }
}  // namespace dev

template <class F>
__global__ void pfor_kernel(long long n, F f) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) f(i);
}
template <class F>
inline void pfor(long long n, F f, stream_t s) {
    if (n <= 0) return;
    long long blocks = (n + 255) / 256;
    if (blocks > 148LL * 16) blocks = 148LL * 16;
    pfor_kernel<F><<<(unsigned)blocks, 256, 0, s>>>(n, f);
    MVD_CUDA_CHECK(cudaGetLastError());
}

#else  // ------------------------------------------------------------------ host emulation (tests only)
typedef void* stream_t;

struct HostExec {
    int nthreads;
    std::vector<double> ss;
    std::vector<float> mm;
    explicit HostExec(int n) : nthreads(n), ss(n, 0.0), mm(n, -1.f) {}
    template <class F> void phase(F&& f) { for (int tid = 0; tid < nthreads; ++tid) f(tid); }
    void stash(int tid, double s, float m) { ss[tid] = s; mm[tid] = m; }
    void unstash(int tid, double& s, float& m) { s = ss[tid]; m = mm[tid]; }
};

namespace dev {
inline void set_device(int) {}
inline void* alloc(size_t n) { void* p = std::malloc(n ? n : 1); if (!p) throw Error("host-emu malloc failed"); return p; }
inline void free_(void* p) { std::free(p); }
inline void h2d(void* d, const void* h, size_t n, stream_t) { std::memcpy(d, h, n); }
inline void d2h(void* h, const void* d, size_t n, stream_t) { std::memcpy(h, d, n); }
inline void d2d(void* d, const void* s_, size_t n, stream_t) { std::memmove(d, s_, n); }
inline void zero(void* d, size_t n, stream_t) { std::memset(d, 0, n); }
inline void sync(stream_t) {}
inline stream_t stream_create() { return nullptr; }
inline stream_t stream_create_high_priority() { return nullptr; }
inline void zero2d(void* d, size_t pitch, size_t width, size_t height, stream_t) {
    for (size_t i = 0; i < height; ++i) std::memset((char*)d + i * pitch, 0, width);
}
inline void stream_destroy(stream_t) {}
typedef void* event_t;
inline event_t event_create() { return nullptr; }
inline void event_destroy(event_t) {}
inline void event_record(event_t, stream_t) {}
inline void stream_wait(stream_t, event_t) {}
}  // namespace dev

template <class F>
inline void pfor(long long n, F f, stream_t) { for (long long i = 0; i < n; ++i) f(i); }
#endif

// --------------------------------------------------------------------------------------------
// per-length kernel registry
// --------------------------------------------------------------------------------------------
struct LenOps {
    int N, R1, R2, R3, T, W, XT, XL;
    int threads, xthreads;
    size_t smem_col, smem_x;
    int ntw_x;                                   // entries of the x-pass stage twiddle tables
    void (*fill_xtw)(cpx* out);                  // host: fills ntw_x entries
    int ntw_col;                                 // entries of the column-pass stage twiddle tables
    void (*fill_ctw)(cpx* out);                  // host: fills ntw_col entries
    void (*launch_col)(int mode, const ColArgs& a, int gx, int gy, stream_t s);
    void (*launch_x)(int kind, const XArgs& a, int nblocks, stream_t s);
};
const LenOps* find_len_ops(int N);                 // nullptr if the length is not instantiated
const std::vector<int>& supported_lengths();       // ascending

}  // namespace mvd
