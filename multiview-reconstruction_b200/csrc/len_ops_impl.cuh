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
// radices above 20 keep up to 30 complex values (plus the kernel spectrum in the convolution stage) in registers: two CTAs per SM
template <class P> constexpr int col_min_blocks() { return min_blocks_for(ColSmem<P>::bytes(), P::THREADS, (P::R1 > 20 || P::R2 > 20 || P::R3 > 20) ? 2 : 3); }
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
// --------------------------------------------------------------------------------------------
// Persistent x-pass kernel (sm_100a): one CTA per resident slot loops over its line groups.  The complex lines of group k+1
// are brought into shared memory by the TMA unit (cp.async.bulk, one bulk copy per line, completion on an mbarrier) while
// group k is transformed, the complex results leave through bulk stores (cp.async.bulk.global.shared::cta, bulk groups), and the
// stage twiddle / twist tables are staged once per CTA instead of once per line group.  Two line buffers alternate.
// --------------------------------------------------------------------------------------------
namespace tma {
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) { while (!mbar_try_wait(bar, parity)) {} }
__device__ __forceinline__ void load_bulk(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void store_bulk(void* dst, const void* src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void wait_group_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void wait_group() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
// L2 prefetch of a contiguous range through the bulk-copy unit: ONE instruction per row instead of one per 128 bytes
__device__ __forceinline__ void prefetch_bulk_l2(const void* src, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
}  // namespace tma

template <class P>
struct XPersist {
    using L = XLay<P>;
    static constexpr bool ok = (L::PAD == 0) && (P::N % 2 == 0);          // unpadded lines, 16-byte multiples per line
    static constexpr size_t bytes() { return sizeof(cpx) * (2 * L::TILE + L::NTAB) + sizeof(LineInfo) * P::XL + 2 * sizeof(unsigned long long); }
};
template <class P> constexpr int xp_min_blocks() { return min_blocks_for(XPersist<P>::bytes(), P::XTHREADS, 4); }

template <class P, int KIND>
__global__ void __launch_bounds__(P::XTHREADS, xp_min_blocks<P>()) x_kernel_p(const XArgs a) {
    extern __shared__ __align__(128) unsigned char mvd_smem[];
    using L = XLay<P>;
    constexpr int XL = P::XL, M = P::N, THREADS = P::XTHREADS;
    constexpr bool HAS_IN = (KIND != X_FWD), HAS_OUT = (KIND == X_FWD || KIND == X_RATIO);
    constexpr unsigned LINE_BYTES = (unsigned)(M * sizeof(cpx));
    cpx* const buf0 = reinterpret_cast<cpx*>(mvd_smem);
    cpx* const buf1 = buf0 + L::TILE;
    cpx* const tabs = buf1 + L::TILE;
    LineInfo* const li = reinterpret_cast<LineInfo*>(tabs + L::NTAB);
    unsigned long long* const bars = reinterpret_cast<unsigned long long*>(li + XL);
    const int tid = (int)threadIdx.x, G = (int)gridDim.x;
    DevExec ex;

    // tables once per CTA
    for (int i = tid; i < L::NTW; i += THREADS) tabs[i] = ld_ro(a.tw + i);
    if (a.xmode == 0 || KIND == X_RATIO || KIND == X_UPDATE) for (int i = tid; i < M; i += THREADS) tabs[L::NTW + i] = ld_ro(a.twist + i);
    if (HAS_IN && tid == 0) { tma::mbar_init(&bars[0], 1); tma::mbar_init(&bars[1], 1); tma::fence_mbar_init(); }
    __syncthreads();

    auto has_work = [&](int bx) -> bool {              // CTA-uniform
        bool any = false;
        const int l0 = a.line0 + bx * XL;
        if constexpr (KIND == X_UPDATE || KIND == X_INV) { for (int i = 0; i < XL; ++i) any = any || x_line_in_box(a, l0 + i); }
        else { for (int i = 0; i < XL; ++i) any = any || x_line_selected(a, l0 + i); }
        return any;
    };
    auto issue_load = [&](int bx, cpx* dst, unsigned long long* bar) {   // one thread
        const int l0 = a.line0 + bx * XL;
        const int n = (a.line_end - l0) < XL ? (a.line_end - l0) : XL;
        tma::mbar_expect_tx(bar, (unsigned)n * LINE_BYTES);
        for (int ln = 0; ln < n; ++ln) tma::load_bulk(dst + ln * L::LS, a.cdata + (long long)(l0 + ln) * a.px, LINE_BYTES, bar);
    };

    int bx = (int)blockIdx.x;
    if (HAS_IN && tid == 0 && bx < a.nblocks && has_work(bx)) issue_load(bx, buf0, &bars[0]);
    unsigned par0 = 0, par1 = 0;
    for (int it = 0; bx < a.nblocks; bx += G, ++it) {
        const int b = it & 1;
        cpx* const sm = b ? buf1 : buf0;
        if (tid == 0) {
            // the other buffer is about to be refilled (by the TMA unit, or by this CTA's stage-1 stores): its bulk store must have
            // finished reading shared memory
            if constexpr (HAS_IN && HAS_OUT) tma::wait_group_read<0>();
            else if constexpr (HAS_OUT) tma::wait_group_read<1>();
            const int nbx = bx + G;
            if (HAS_IN && nbx < a.nblocks && has_work(nbx)) issue_load(nbx, b ? buf0 : buf1, &bars[b ^ 1]);
        }
        if (!has_work(bx)) {
            if constexpr (KIND == X_UPDATE) { if (tid == 0) { a.part_sum[bx] = 0.0; a.part_max[bx] = -1.f; } }
            continue;
        }
        if constexpr (HAS_IN) {
            if (b) { tma::mbar_wait(&bars[1], par1); par1 ^= 1; } else { tma::mbar_wait(&bars[0], par0); par0 ^= 1; }
        }
        x_pass_body<P, KIND, DevExec, true>(ex, a, bx, sm, li, tabs);
        if constexpr (HAS_OUT) {
            tma::fence_proxy_async();                  // generic-proxy writes of the last stage -> visible to the bulk store
            __syncthreads();
            if (tid == 0) {
                const int l0 = a.line0 + bx * XL;
                const int n = (a.line_end - l0) < XL ? (a.line_end - l0) : XL;
                for (int ln = 0; ln < n; ++ln)
                    if (x_line_selected(a, l0 + ln)) tma::store_bulk(a.cdata + (long long)(l0 + ln) * a.px, sm + ln * L::LS, LINE_BYTES);
                tma::commit_group();
            }
        }
    }
    if (HAS_OUT && tid == 0) tma::wait_group<0>();
}

// --------------------------------------------------------------------------------------------
// Warp-autonomous persistent x-pass kernel (sm_100a): ONE WARP PER LINE.  With XT = 32 lanes along a line every stage of a line's
// transform can be dealt to the line's own warp, so the phases of a line are separated by __syncwarp() instead of CTA barriers and
// every warp is an independent pipeline: its own pair of line buffers, its own mbarriers, its own bulk loads (cp.async.bulk ->
// mbarrier complete_tx) and bulk stores.  Nothing waits for the slowest warp of a CTA any more (the barrier stall was the top stall
// of the CTA-wide version: 2.3 - 3.5 cycles per issue); the price is the last-stage butterflies running on N / RL of 32 lanes.
// The CTA only shares the stage twiddle / twist tables.
// --------------------------------------------------------------------------------------------
struct WarpExec {
    double s_;
    float m_;
    template <class F> __device__ __forceinline__ void phase(F&& f) { f((int)(threadIdx.x & 31)); __syncwarp(); }
    __device__ __forceinline__ void stash(int, double s, float m) { s_ = s; m_ = m; }
    __device__ __forceinline__ void unstash(int, double& s, float& m) { s = s_; m = m_; }
};
template <class P, int WPC>
struct XWarp {
    using L = XLay<P>;
    // worth it when the first-stage butterflies fill the warp (N / R1 of 32 lanes; measured: N = 540 with 30 lanes -3 % on the quotient
    // pass, N = 288 with 18 lanes +35 %)
    static constexpr bool ok = (L::PAD == 0) && (P::N % 2 == 0) && P::XL == 1 && P::XT == 32 && (P::N / P::R1 >= 28) && (P::N / P::R1 <= 32);
    static constexpr size_t bytes() {
        return sizeof(cpx) * (size_t)(2 * WPC * L::LS + L::NTAB) + sizeof(LineInfo) * WPC + 2 * WPC * sizeof(unsigned long long);
    }
};
template <class P, int WPC> constexpr int xw_min_blocks() { return min_blocks_for(XWarp<P, WPC>::bytes(), 32 * WPC, 16 / WPC); }

template <class P, int KIND, int WPC>
__global__ void __launch_bounds__(32 * WPC, xw_min_blocks<P, WPC>()) x_kernel_w(const XArgs a) {
    extern __shared__ __align__(128) unsigned char mvd_smem[];
    using L = XLay<P>;
    constexpr int M = P::N, THREADS = 32 * WPC;
    constexpr bool HAS_IN = (KIND != X_FWD), HAS_OUT = (KIND == X_FWD || KIND == X_RATIO);
    constexpr unsigned LINE_BYTES = (unsigned)(M * sizeof(cpx));
    const int warp = (int)threadIdx.x >> 5, lane = (int)threadIdx.x & 31;
    cpx* const bufs = reinterpret_cast<cpx*>(mvd_smem);                       // [WPC][2][LS]
    cpx* const tabs = bufs + 2 * WPC * L::LS;
    LineInfo* const li_all = reinterpret_cast<LineInfo*>(tabs + L::NTAB);
    unsigned long long* const bars = reinterpret_cast<unsigned long long*>(li_all + WPC) + 2 * warp;
    LineInfo* const li = li_all + warp;
    cpx* const buf0 = bufs + 2 * warp * L::LS;
    cpx* const buf1 = buf0 + L::LS;

    for (int i = (int)threadIdx.x; i < L::NTW; i += THREADS) tabs[i] = ld_ro(a.tw + i);
    if (a.xmode == 0 || KIND == X_RATIO || KIND == X_UPDATE) for (int i = (int)threadIdx.x; i < M; i += THREADS) tabs[L::NTW + i] = ld_ro(a.twist + i);
    if (HAS_IN && lane == 0) { tma::mbar_init(&bars[0], 1); tma::mbar_init(&bars[1], 1); tma::fence_mbar_init(); }
    __syncthreads();

    // work items: the lines of the launch, or (nrect > 0) the lines of the selection rectangles; item k -> line index relative to line0
    const int nlines = a.nrect > 0 ? a.rect_start[a.nrect] : a.line_end - a.line0;
    const int stride = (int)gridDim.x * WPC;
    auto line_of = [&](int k) -> int {
        if (a.nrect <= 0) return k;
        int r = 0;
        while (r + 1 < a.nrect && k >= a.rect_start[r + 1]) ++r;
        const int kk = k - a.rect_start[r], w = a.rect[r][1] - a.rect[r][0];
        return (a.rect[r][0] + kk % w) + a.ty * (a.rect[r][2] + kk / w) - a.line0;
    };
    auto issue_load = [&](int ln, cpx* dst, unsigned long long* bar) {   // one lane
        tma::mbar_expect_tx(bar, LINE_BYTES);
        tma::load_bulk(dst, a.cdata + (long long)(a.line0 + ln) * a.px, LINE_BYTES, bar);
    };
    int k = (int)blockIdx.x * WPC + warp;
    unsigned par0 = 0, par1 = 0;
    WarpExec ex;
    // x range of the real rows this tile reads, as a 16-byte aligned byte range (bulk prefetch granularity); rows start 16-byte aligned
    // when the local x extent is a multiple of 4
    int pf_x0 = clampi(a.org[0] - a.goff[0], 0, a.vol[0] - 1) & ~3;
    int pf_x1 = clampi(a.org[0] - a.goff[0] + 2 * M, 0, a.vol[0]);
    const unsigned pf_bytes = (unsigned)((pf_x1 - pf_x0) * 4) & ~15u;
    const bool pf_rows = a.pf_dist > 0 && (a.vol[0] & 3) == 0 && ((reinterpret_cast<unsigned long long>(a.src) & 15ull) == 0) && pf_bytes > 0;
    // lines outside the responsibility box of the update / inverse pass, or rejected by the line filter, are skipped (no load, no work);
    // a rectangle list already holds the selected lines only
    const bool filtered = a.nrect <= 0 && ((KIND == X_UPDATE || KIND == X_INV) || a.fin[1] > a.fin[0] || a.fout[1] > a.fout[0]);
    auto has_work = [&](int kk) -> bool {
        if constexpr (KIND == X_UPDATE || KIND == X_INV) return x_line_in_box(a, a.line0 + kk);
        else return x_line_selected(a, a.line0 + kk);
    };
    if (filtered) {
        while (k < nlines && !has_work(k)) k += stride;
    }
    int ln = k < nlines ? line_of(k) : 0;
    if (HAS_IN && lane == 0 && k < nlines) issue_load(ln, buf0, &bars[0]);
    double wsum = 0.0;
    float wmax = -1.f;
    for (int it = 0; k < nlines; ++it) {
        const int b = it & 1;
        cpx* const sm = b ? buf1 : buf0;
        int nk = k + stride;
        if (filtered) {
            while (nk < nlines && !has_work(nk)) nk += stride;
        }
        const bool more = nk < nlines;
        const int nl = more ? line_of(nk) : 0;
        // the pass stores into `sm` from its first phase on: the bulk store that last read this buffer (two lines ago) must be done
        if constexpr (HAS_OUT && !HAS_IN) { if (lane == 0) tma::wait_group_read<1>(); __syncwarp(); }
        if constexpr (HAS_IN) {
            if (b) { tma::mbar_wait(&bars[1], par1); par1 ^= 1; } else { tma::mbar_wait(&bars[0], par0); par0 ^= 1; }
        }
        auto prefetch_next = [&]() {
            if constexpr (HAS_IN) {
                if (lane == 0) {
                    if constexpr (HAS_OUT) tma::wait_group_read<0>();     // the other buffer's bulk store (previous line) has been read out
                    if (more) {
                        issue_load(nl, b ? buf0 : buf1, &bars[b ^ 1]);
                        if constexpr (KIND == X_RATIO || KIND == X_UPDATE) {
                            // observed-image row of that line -> L2 (one bulk prefetch; rows outside the volume carry no data)
                            const int l = a.line0 + nl;
                            const int gy = a.org[1] + l % a.ty, gz = a.org[2] + l / a.ty;
                            if (pf_rows && (unsigned)gy < (unsigned)a.gdim[1] && (unsigned)gz < (unsigned)a.gdim[2]) {
                                const int ly = clampi(gy - a.goff[1], 0, a.vol[1] - 1), lz = clampi(gz - a.goff[2], 0, a.vol[2] - 1);
                                const float* p = a.src + ((long long)lz * a.vol[1] + ly) * (long long)a.vol[0] + pf_x0;
                                tma::prefetch_bulk_l2(p, pf_bytes);
                                if constexpr (KIND == X_UPDATE) tma::prefetch_bulk_l2(a.weight + (p - a.src), pf_bytes);
                            }
                        }
                    }
                }
            }
        };
        x_pass_body<P, KIND, WarpExec, true>(ex, a, ln, sm, li, tabs, prefetch_next);
        if constexpr (HAS_OUT) {
            tma::fence_proxy_async();                  // generic-proxy writes of the last stage -> visible to the bulk store
            __syncwarp();
            if (lane == 0) {
                tma::store_bulk(a.cdata + (long long)(a.line0 + ln) * a.px, sm, LINE_BYTES);
                tma::commit_group();
            }
        }
        if constexpr (KIND == X_UPDATE) {              // signed change statistics of this lane's voxels, kept in registers across the lines
            double s; float m;
            ex.unstash(lane, s, m);
            wsum += s; wmax = m > wmax ? m : wmax;
        }
        k = nk;
        ln = nl;
    }
    if (HAS_OUT && lane == 0) tma::wait_group<0>();
    if constexpr (KIND == X_UPDATE) {
        // warp-shuffle reduction (fixed tree: deterministic), one partial per warp; the remaining slots of the launch are neutral
        for (int o = 16; o > 0; o >>= 1) {
            wsum += __shfl_down_sync(0xffffffffu, wsum, o);
            const float om = __shfl_down_sync(0xffffffffu, wmax, o);
            wmax = om > wmax ? om : wmax;
        }
        const int gw = (int)blockIdx.x * WPC + warp;
        if (lane == 0) { a.part_sum[gw] = wsum; a.part_max[gw] = wmax; }
        for (int i = stride + (int)blockIdx.x * THREADS + (int)threadIdx.x; i < a.nblocks; i += (int)gridDim.x * THREADS) { a.part_sum[i] = 0.0; a.part_max[i] = -1.f; }
    }
}

// cudaFuncSetAttribute is a per-device setting: remember which devices of this process already have it (the reference drives several
// devices from one process, one Java thread each -- MultiViewDeconvolutionSeq.java:92-150)
inline bool first_use_on_current_device(std::atomic<unsigned long long>& mask) {
    int d = 0;
    MVD_CUDA_CHECK(cudaGetDevice(&d));
    const unsigned long long bit = 1ull << (d & 63);
    return (mask.fetch_or(bit) & bit) == 0;
}
#endif

inline bool x_persistent_enabled() {   // MVD_XPERSIST=0 selects the one-line-group-per-CTA kernels (A/B measurements)
    static const bool v = [] { const char* e = std::getenv("MVD_XPERSIST"); return !(e && std::atoi(e) == 0); }();
    return v;
}
inline int carveout_pref() {   // MVD_CARVEOUT: -1 = driver default, 0..100 = preferred shared-memory carveout in percent
    static const int v = [] { const char* e = std::getenv("MVD_CARVEOUT"); return e ? std::atoi(e) : -1; }();
    return v;
}

// P: plan of the column (y/z) kernels; PX: plan of the x kernels (same length, its own radices / threading).  The two may differ
// because the scrambled frequency order of an axis only has to be consistent among the kernels that transform that axis.
// PXP: the x plan of the persistent kernels -- PX with its own number of lines per group (smaller groups = more resident CTAs, i.e. more
// independent barrier domains per SM; measured on c3: forward / quotient pass -5 / -7 % with 4-line groups, update pass +16 %)
inline int x_warp_mode() {   // MVD_XWARP=0 selects the CTA-wide persistent kernels (A/B measurements)
    static const int v = [] { const char* e = std::getenv("MVD_XWARP"); return e ? std::atoi(e) : 1; }();
    return v;
}

template <class P, class PX, class PXP>
struct LenImpl {
    using PXW = Plan<PX::N, PX::R1, PX::R2, PX::R3, 1, 1, 32, 1>;      // one warp per line
    static constexpr int WPC = 4;                                      // warps per CTA of the warp-autonomous kernels
    static_assert(P::N == PX::N, "column and x plans must describe the same length");
    static_assert(PXP::N == PX::N && PXP::R1 == PX::R1 && PXP::R2 == PX::R2 && PXP::R3 == PX::R3 && PXP::XT == PX::XT, "persistent x plan = x plan with another group size");
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
        // persistent TMA-staged kernels for the passes that write a complex tile (measured: forward -18 %, quotient -10 %); the update /
        // inverse passes are bound by their real-space loads and keep one line group per CTA (three resident CTAs instead of two)
        if constexpr (XWarp<PXW, WPC>::ok) {
            // MVD_XWARP: 1 (default) = forward / quotient passes; 2 = update / inverse passes too (measured on c3: update pass 0.91 ms against
            // 0.83 ms of the one-shot kernel -- 124 registers leave 16 warps per SM where the one-shot kernel has 24 to hide its row loads)
            if (x_persistent_enabled() && (x_warp_mode() >= 2 || (x_warp_mode() == 1 && (KIND == X_FWD || KIND == X_RATIO)))) {
                if (xp_warp<KIND>(a, nblocks, s)) return;
            }
        }
        if constexpr (XPersist<PXP>::ok && (KIND == X_FWD || KIND == X_RATIO)) {
            if (x_persistent_enabled()) { xp_persistent<KIND>(a, (a.line_end - a.line0 + PXP::XL - 1) / PXP::XL, s); return; }
        }
        static std::atomic<unsigned long long> attr_mask{0};
        if (first_use_on_current_device(attr_mask)) {
            MVD_CUDA_CHECK(cudaFuncSetAttribute(x_kernel<PX, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_x));
            if (carveout_pref() >= 0) MVD_CUDA_CHECK(cudaFuncSetAttribute(x_kernel<PX, KIND>, cudaFuncAttributePreferredSharedMemoryCarveout, carveout_pref()));
        }
        x_kernel<PX, KIND><<<nblocks, PX::XTHREADS, smem_x, s>>>(a);
        MVD_CUDA_CHECK(cudaGetLastError());
#endif
    }
#ifndef MVD_HOST_EMU
    // warp-autonomous launch: resident CTAs x WPC warps, every warp strides over the lines of the launch
    template <int KIND>
    static bool xp_warp(XArgs a, int nblocks, stream_t s) {
        constexpr size_t smem = XWarp<PXW, WPC>::bytes();
        static std::atomic<unsigned long long> attr_mask{0};
        static std::atomic<int> slots[64];
        int d = 0;
        MVD_CUDA_CHECK(cudaGetDevice(&d));
        if (first_use_on_current_device(attr_mask)) {
            MVD_CUDA_CHECK(cudaFuncSetAttribute(x_kernel_w<PXW, KIND, WPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int per_sm = 0, sms = 0;
            MVD_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, x_kernel_w<PXW, KIND, WPC>, 32 * WPC, smem));
            MVD_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, d));
            slots[d & 63] = (per_sm > 0 ? per_sm : 1) * sms;
        }
        const int nlines = a.nrect > 0 ? a.rect_start[a.nrect] : a.line_end - a.line0;
        int grid = slots[d & 63].load();
        if (grid <= 0) grid = 148;
        if (grid > (nlines + WPC - 1) / WPC) grid = (nlines + WPC - 1) / WPC;
        if (grid < 1) return true;
        if (KIND == X_UPDATE && grid * WPC > nblocks) return false;      // one statistics slot per warp
        a.nblocks = nblocks;
        // bulk L2 prefetch of the real rows of the line a warp handles next (the forward pass is faster without)
        a.pf_dist = (a.pf_dist > 0 && KIND != X_FWD) ? grid * WPC : 0;
        x_kernel_w<PXW, KIND, WPC><<<grid, 32 * WPC, smem, s>>>(a);
        MVD_CUDA_CHECK(cudaGetLastError());
        return true;
    }
    // persistent launch: one CTA per resident slot of the device (occupancy query on first use), pf_dist = grid size so that the
    // software L2 prefetch of the real rows targets the line group this CTA handles next
    template <int KIND>
    static void xp_persistent(XArgs a, int nblocks, stream_t s) {
        constexpr size_t smem = XPersist<PXP>::bytes();
        static std::atomic<unsigned long long> attr_mask{0};
        static std::atomic<int> slots[64];
        int d = 0;
        MVD_CUDA_CHECK(cudaGetDevice(&d));
        if (first_use_on_current_device(attr_mask)) {
            MVD_CUDA_CHECK(cudaFuncSetAttribute(x_kernel_p<PXP, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int per_sm = 0, sms = 0;
            MVD_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, x_kernel_p<PXP, KIND>, PXP::XTHREADS, smem));
            MVD_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, d));
            slots[d & 63] = (per_sm > 0 ? per_sm : 1) * sms;
        }
        int grid = slots[d & 63].load();
        if (grid <= 0) grid = 148;
        if (grid > nblocks) grid = nblocks;
        a.pf_dist = (a.pf_dist > 0 && KIND != X_FWD) ? grid : 0;     // the forward pass is faster without the software prefetch (measured)
        a.nblocks = nblocks;
        x_kernel_p<PXP, KIND><<<grid, PXP::XTHREADS, smem, s>>>(a);
        MVD_CUDA_CHECK(cudaGetLastError());
    }
#endif
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

#define MVD_DEFINE_LEN(N, R1, R2, R3, T, W, XR1, XR2, XR3, XT, XL, XLP) \
    namespace mvd { const LenOps* len_ops_##N() { return LenImpl<Plan<N, R1, R2, R3, T, W, 1, 1>, Plan<N, XR1, XR2, XR3, 1, 1, XT, XL>, Plan<N, XR1, XR2, XR3, 1, 1, XT, XLP>>::ops(); } }
