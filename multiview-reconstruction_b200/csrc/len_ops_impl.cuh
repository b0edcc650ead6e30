// Instantiates the seven pass kernels for one FFT length and exposes them as a LenOps record.
// Included by the generated per-length translation units (gen/len_<N>.cu) so that lengths compile in parallel.
#pragma once
#include "backend.h"

namespace mvd {

template <class P>
constexpr int plan_min_blocks() {
    // aim for 3 resident CTAs per SM when shared memory and the 2048-thread limit allow it
    constexpr size_t smem = sizeof(cpx) * P::N * (P::W + 1) + 1024;
    int by_smem = int((227u * 1024u) / smem);
    int by_thr = 2048 / P::THREADS;
    int b = by_smem < by_thr ? by_smem : by_thr;
    if (b > 3) b = 3;
    // do not ask for fewer than 64 registers per thread
    while (b > 1 && 65536 / (P::THREADS * b) < 64) --b;
    return b < 1 ? 1 : b;
}

#ifndef MVD_HOST_EMU
template <class P, int MODE>
__global__ void __launch_bounds__(P::THREADS, plan_min_blocks<P>()) col_kernel(const ColArgs a) {
    extern __shared__ __align__(16) unsigned char mvd_smem[];
    DevExec ex;
    col_pass_body<P, MODE>(ex, a, (int)blockIdx.x, (int)blockIdx.y, reinterpret_cast<cpx*>(mvd_smem));
}
template <class P, int KIND>
__global__ void __launch_bounds__(P::THREADS, plan_min_blocks<P>()) x_kernel(const XArgs a) {
    extern __shared__ __align__(16) unsigned char mvd_smem[];
    DevExec ex;
    cpx* sm = reinterpret_cast<cpx*>(mvd_smem);
    LineInfo* li = reinterpret_cast<LineInfo*>(sm + XSmem<P>::TILE);
    x_pass_body<P, KIND>(ex, a, (int)blockIdx.x, sm, li);
}
#endif

template <class P>
struct LenImpl {
    static constexpr size_t smem_col = sizeof(cpx) * P::N * P::W;
    static constexpr size_t smem_x = XSmem<P>::bytes() > sizeof(double) * P::THREADS + sizeof(float) * P::THREADS
                                         ? XSmem<P>::bytes()
                                         : sizeof(double) * P::THREADS + sizeof(float) * P::THREADS + sizeof(LineInfo) * P::W;

    template <int MODE>
    static void col(const ColArgs& a, int gx, int gy, stream_t s) {
#ifdef MVD_HOST_EMU
        (void)s;
        std::vector<cpx> sm(P::N * P::W);
        HostExec ex(P::THREADS);
        for (int by = 0; by < gy; ++by)
            for (int bx = 0; bx < gx; ++bx) col_pass_body<P, MODE>(ex, a, bx, by, sm.data());
#else
        static bool attr_done = false;
        if (!attr_done) {
            MVD_CUDA_CHECK(cudaFuncSetAttribute(col_kernel<P, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_col));
            attr_done = true;
        }
        col_kernel<P, MODE><<<dim3(gx, gy), P::THREADS, smem_col, s>>>(a);
        MVD_CUDA_CHECK(cudaGetLastError());
#endif
    }
    template <int KIND>
    static void xp(const XArgs& a, int nblocks, stream_t s) {
#ifdef MVD_HOST_EMU
        (void)s;
        std::vector<unsigned char> raw(smem_x + 16);
        cpx* sm = reinterpret_cast<cpx*>(raw.data());
        LineInfo* li = reinterpret_cast<LineInfo*>(sm + XSmem<P>::TILE);
        HostExec ex(P::THREADS);
        for (int bx = 0; bx < nblocks; ++bx) x_pass_body<P, KIND>(ex, a, bx, sm, li);
#else
        static bool attr_done = false;
        if (!attr_done) {
            MVD_CUDA_CHECK(cudaFuncSetAttribute(x_kernel<P, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_x));
            attr_done = true;
        }
        x_kernel<P, KIND><<<nblocks, P::THREADS, smem_x, s>>>(a);
        MVD_CUDA_CHECK(cudaGetLastError());
#endif
    }
    static void launch_col(int mode, const ColArgs& a, int gx, int gy, stream_t s) {
        switch (mode) {
            case COL_FWD: col<COL_FWD>(a, gx, gy, s); break;
            case COL_INV: col<COL_INV>(a, gx, gy, s); break;
            case COL_CONV: col<COL_CONV>(a, gx, gy, s); break;
            default: throw Error("bad column mode");
        }
    }
    static void launch_x(int kind, const XArgs& a, int nblocks, stream_t s) {
        switch (kind) {
            case X_FWD: xp<X_FWD>(a, nblocks, s); break;
            case X_RATIO: xp<X_RATIO>(a, nblocks, s); break;
            case X_UPDATE: xp<X_UPDATE>(a, nblocks, s); break;
            case X_INV: xp<X_INV>(a, nblocks, s); break;
            default: throw Error("bad x-pass kind");
        }
    }
    static const LenOps* ops() {
        static const LenOps o = {P::N, P::R1, P::R2, P::R3, P::T, P::W, P::THREADS, smem_col, smem_x, &launch_col, &launch_x};
        return &o;
    }
};

}  // namespace mvd

#define MVD_DEFINE_LEN(N, R1, R2, R3, T, W) \
    namespace mvd { const LenOps* len_ops_##N() { return LenImpl<Plan<N, R1, R2, R3, T, W>>::ops(); } }
