// Instantiates the seven pass kernels for one FFT length and exposes them as a LenOps record.
// Included by the generated per-length translation units (gen/len_<N>.cu) so that lengths compile in parallel.
#pragma once
#include <atomic>

#include "backend.h"

namespace mvd {

constexpr int min_blocks_for(size_t smem_bytes, int threads, int want) {
    // resident CTAs per SM we ask the compiler to make room for: bounded by shared memory, the 2048-thread limit and a
    // floor of 64 registers per thread
    int by_smem = int((227u * 1024u) / (smem_bytes + 1024));
    int by_thr = 2048 / threads;
    int b = by_smem < by_thr ? by_smem : by_thr;
    if (b > want) b = want;
#ifndef MVD_MIN_REGS
#define MVD_MIN_REGS 64
#endif
    while (b > 1 && 65536 / (threads * b) < MVD_MIN_REGS) --b;
    return b < 1 ? 1 : b;
}
template <class P> constexpr int col_min_blocks() { return min_blocks_for(ColSmem<P>::bytes(), P::THREADS, 3); }
template <class P> constexpr int x_min_blocks() { return min_blocks_for(XLay<P>::bytes(), P::XTHREADS, (P::R1 > 20 || P::R2 > 20 || P::R3 > 20) ? 3 : 4); }

#ifndef MVD_HOST_EMU
template <class P, int MODE>
__global__ void __launch_bounds__(P::THREADS, col_min_blocks<P>()) col_kernel(const ColArgs a) {
    extern __shared__ __align__(16) unsigned char mvd_smem[];
    DevExec ex;
    col_pass_body<P, MODE>(ex, a, (int)blockIdx.x, (int)blockIdx.y, reinterpret_cast<cpx*>(mvd_smem));
}
template <class P, int KIND>
__global__ void __launch_bounds__(P::XTHREADS, x_min_blocks<P>()) x_kernel(const XArgs a) {
    extern __shared__ __align__(16) unsigned char mvd_smem[];
    DevExec ex;
    cpx* sm = reinterpret_cast<cpx*>(mvd_smem);
    LineInfo* li = reinterpret_cast<LineInfo*>(sm + XLay<P>::TILE + XLay<P>::NTAB);
    x_pass_body<P, KIND>(ex, a, (int)blockIdx.x, sm, li);
}
#endif

#ifndef MVD_HOST_EMU
// cudaFuncSetAttribute is a per-device setting: remember which devices of this process already have it (the reference drives several
// devices from one process, one Java thread each -- MultiViewDeconvolutionSeq.java:92-150)
inline bool first_use_on_current_device(std::atomic<unsigned long long>& mask) {
    int d = 0;
    MVD_CUDA_CHECK(cudaGetDevice(&d));
    const unsigned long long bit = 1ull << (d & 63);
    return (mask.fetch_or(bit) & bit) == 0;
}
#endif

inline int carveout_pref() {   // MVD_CARVEOUT: -1 = driver default, 0..100 = preferred shared-memory carveout in percent
    static const int v = [] { const char* e = std::getenv("MVD_CARVEOUT"); return e ? std::atoi(e) : -1; }();
    return v;
}

// P: plan of the column (y/z) kernels; PX: plan of the x kernels (same length, its own radices / threading).  The two may differ
// because the scrambled frequency order of an axis only has to be consistent among the kernels that transform that axis.
template <class P, class PX>
struct LenImpl {
    static_assert(P::N == PX::N, "column and x plans must describe the same length");
    static constexpr size_t smem_col = ColSmem<P>::bytes();
    static constexpr size_t smem_x = XLay<PX>::bytes();
    static_assert(sizeof(cpx) * XLay<PX>::TILE >= (sizeof(double) + sizeof(float)) * PX::XTHREADS, "reduction scratch must fit in the tile");

    template <int MODE>
    static void col(const ColArgs& a, int gx, int gy, stream_t s) {
#ifdef MVD_HOST_EMU
        (void)s;
        std::vector<cpx> sm(smem_col / sizeof(cpx) + 1);
        HostExec ex(P::THREADS);
        for (int by = 0; by < gy; ++by)
            for (int bx = 0; bx < gx; ++bx) col_pass_body<P, MODE>(ex, a, bx, by, sm.data());
#else
        static std::atomic<unsigned long long> attr_mask{0};
        if (first_use_on_current_device(attr_mask)) {
            MVD_CUDA_CHECK(cudaFuncSetAttribute(col_kernel<P, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_col));
            if (carveout_pref() >= 0) MVD_CUDA_CHECK(cudaFuncSetAttribute(col_kernel<P, MODE>, cudaFuncAttributePreferredSharedMemoryCarveout, carveout_pref()));
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
        LineInfo* li = reinterpret_cast<LineInfo*>(sm + XLay<PX>::TILE + XLay<PX>::NTAB);
        HostExec ex(PX::XTHREADS);
        for (int bx = 0; bx < nblocks; ++bx) x_pass_body<PX, KIND>(ex, a, bx, sm, li);
#else
        static std::atomic<unsigned long long> attr_mask{0};
        if (first_use_on_current_device(attr_mask)) {
            MVD_CUDA_CHECK(cudaFuncSetAttribute(x_kernel<PX, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_x));
            if (carveout_pref() >= 0) MVD_CUDA_CHECK(cudaFuncSetAttribute(x_kernel<PX, KIND>, cudaFuncAttributePreferredSharedMemoryCarveout, carveout_pref()));
        }
        x_kernel<PX, KIND><<<nblocks, PX::XTHREADS, smem_x, s>>>(a);
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
        static const LenOps o = {P::N, P::R1, P::R2, P::R3, P::T, P::W, PX::XT, PX::XL, P::THREADS, PX::XTHREADS, smem_col, smem_x,
                                 XLay<PX>::NTW, &fill_xtw<PX>, StageTw<P>::NTW, &fill_stage_tw<P>, &launch_col, &launch_x};
        return &o;
    }
};

}  // namespace mvd

#define MVD_DEFINE_LEN(N, R1, R2, R3, T, W, XR1, XR2, XR3, XT, XL) \
    namespace mvd { const LenOps* len_ops_##N() { return LenImpl<Plan<N, R1, R2, R3, T, W, 1, 1>, Plan<N, XR1, XR2, XR3, 1, 1, XT, XL>>::ops(); } }
