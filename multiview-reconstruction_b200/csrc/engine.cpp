// Engine implementation.  Compiled by nvcc (-x cu) for the product and by g++ with -DMVD_HOST_EMU for the
// CPU-side unit tests of the index math (tests/host).
#include "engine.h"

#include <algorithm>
#include <cmath>

namespace mvd {

// ------------------------------------------------------------------------------------------------
// twiddle tables
// ------------------------------------------------------------------------------------------------
Tables::~Tables() {
    for (auto& kv : tw_) dev::free_(kv.second);
    for (auto& kv : twist_) dev::free_(kv.second);
    for (auto& kv : xtw_) dev::free_(kv.second);
}
const cpx* Tables::xtw(int M) {
    auto it = xtw_.find(M);
    if (it != xtw_.end()) return it->second;
    const LenOps* o = find_len_ops(M);
    if (!o) throw Error("no kernels for this length");
    std::vector<cpx> h((size_t)o->ntw_x + 1);
    o->fill_xtw(h.data());
    cpx* d = (cpx*)dev::alloc(sizeof(cpx) * h.size());
    dev::h2d(d, h.data(), sizeof(cpx) * h.size(), stream_);
    dev::sync(stream_);
    xtw_[M] = d;
    return d;
}
const cpx* Tables::tw(int N) {          // per-position stage twiddles of the column passes (LenOps::fill_ctw)
    auto it = tw_.find(N);
    if (it != tw_.end()) return it->second;
    const LenOps* o = find_len_ops(N);
    if (!o) throw Error("no kernels for this length");
    std::vector<cpx> h((size_t)o->ntw_col + 1);
    o->fill_ctw(h.data());
    cpx* d = (cpx*)dev::alloc(sizeof(cpx) * h.size());
    dev::h2d(d, h.data(), sizeof(cpx) * h.size(), stream_);
    dev::sync(stream_);
    tw_[N] = d;
    return d;
}
const cpx* Tables::twist(int M) {
    auto it = twist_.find(M);
    if (it != twist_.end()) return it->second;
    std::vector<cpx> h(M);
    for (int m = 0; m < M; ++m) {
        const double a = M_PI * (double)m / (double)(2 * M);
        h[m] = cpx{(float)std::cos(a), (float)-std::sin(a)};
    }
    cpx* d = (cpx*)dev::alloc(sizeof(cpx) * M);
    dev::h2d(d, h.data(), sizeof(cpx) * M, stream_);
    dev::sync(stream_);
    twist_[M] = d;
    return d;
}

// ------------------------------------------------------------------------------------------------
// tile planner.  Same validity rule as the reference's halo'd blocks (BlockGeneratorFixedSizePrecise.java:68-91,
// DeconView.java:155-157): a tile's output is trusted where neither of the two chained convolutions read across the
// tile edge.  Outside the global volume the quotient is known to be 1 (no image data), which is why tiles that
// contain a volume boundary only need the single-convolution margin there.
// ------------------------------------------------------------------------------------------------
AxisTiling plan_axis(int gdim, int a, int b, Reach r1, Reach r2, bool is_x, int max_len, bool two_exchanges) {
    const int Lsum = r1.lo + r2.lo, Rsum = r1.hi + r2.hi;
    const int Lmax = std::max(r1.lo, r2.lo), Rmax = std::max(r1.hi, r2.hi);
    AxisTiling best;
    static const int force_x = [] { const char* e = std::getenv("MVD_FORCE_XLEN"); return e ? std::atoi(e) : 0; }();       // experiments only
    static const int exclude_x = [] { const char* e = std::getenv("MVD_EXCLUDE_XLEN"); return e ? std::atoi(e) : 0; }();   // experiments only
    for (int len : supported_lengths()) {
        if (len > max_len) break;
        if (is_x && len == exclude_x) continue;
        if (is_x && force_x && len != force_x) continue;
        const int T = is_x ? 2 * len : len;
        AxisTiling cur;
        cur.T = T;
        int pos = a;
        // a shard boundary behaves like a volume face when the quotient is exchanged too (scheme B): one reach is enough
        int o = (a == 0 || two_exchanges) ? a - Lmax : a - Lsum;
        bool ok = true;
        while (pos < b) {
            int vend = o + T - Rsum;
            if ((b == gdim || two_exchanges) && o + T >= b + Rmax) vend = b;
            const int hi = std::min(vend, b);
            if (hi <= pos || cur.tiles.size() > 65536) { ok = false; break; }
            cur.tiles.push_back(AxisTile{o, pos, hi});
            pos = hi;
            o = pos - Lsum;
        }
        if (!ok) continue;
        if (best.T == 0 || cur.cost() < best.cost()) best = cur;
    }
    if (best.T == 0) throw Error("no supported FFT length fits this axis (volume + PSF too large for max_len?)");
    return best;
}

// ------------------------------------------------------------------------------------------------
// small device functors
// ------------------------------------------------------------------------------------------------
struct PlaceKernel {   // scatter the (tiny) kernel into a zeroed tile with its centre at the origin
    const float* k;
    float* out;
    int kd0, kd1, kd2, T0, T1, T2;
    float scale;
    int negacyclic;
    MVD_HD void operator()(long long i) const {
        const int jx = (int)(i % kd0), jy = (int)((i / kd0) % kd1), jz = (int)(i / ((long long)kd0 * kd1));
        int px = jx - kd0 / 2, py = jy - kd1 / 2, pz = jz - kd2 / 2;
        float s = scale;
        if (px < 0) { px += T0; if (negacyclic) s = -s; }     // skew-circular wrap in x
        if (py < 0) py += T1;
        if (pz < 0) pz += T2;
        out[((long long)pz * T1 + py) * T0 + px] = s * k[i];
    }
};
struct FillParts {      // unused partial-statistics slots must read as {0, -1}
    double* ps; float* pm;
    MVD_HD void operator()(long long i) const { ps[i] = 0.0; pm[i] = -1.f; }
};
// Deterministic two-level reduction: lane l adds the partials l, l + 256, ... in index order, then one thread adds the 256 lane sums in
// order.  The loads are issued in batches of 8 ahead of the (ordered) additions so the single CTA is not bound by one L2 latency per term.
struct ReduceParts1 {
    const double* ps; const float* pm; int n; double* ts; float* tm;
    MVD_HD void operator()(long long lane) const {
        double s = 0.0; float m = -1.f;
        int i = (int)lane;
        for (; i + 7 * 256 < n; i += 8 * 256) {
            double a[8]; float b[8];
            for (int k = 0; k < 8; ++k) { a[k] = ps[i + k * 256]; b[k] = pm[i + k * 256]; }
            for (int k = 0; k < 8; ++k) { s += a[k]; m = b[k] > m ? b[k] : m; }
        }
        for (; i < n; i += 256) { s += ps[i]; m = pm[i] > m ? pm[i] : m; }
        ts[lane] = s; tm[lane] = m;
    }
};
struct ReduceParts2 {
    const double* ts; const float* tm; double* out;
    MVD_HD void operator()(long long) const {
        double s = 0.0; float m = -1.f;
        for (int i = 0; i < 256; i += 16) {
            double a[16]; float b[16];
            for (int k = 0; k < 16; ++k) { a[k] = ts[i + k]; b[k] = tm[i + k]; }
            for (int k = 0; k < 16; ++k) { s += a[k]; m = b[k] > m ? b[k] : m; }
        }
        out[0] = s; out[1] = (double)m;
    }
};

#ifndef MVD_HOST_EMU
// ReduceParts1 + ReduceParts2 in one launch: the 256 lane sums go through shared memory instead of a second kernel (same order of
// additions, hence the same bits; one launch and ~20 us less behind every view update)
__global__ void __launch_bounds__(256) reduce_parts_kernel(ReduceParts1 r1, double* out) {
    __shared__ double ss[256];
    __shared__ float sm[256];
    r1.ts = ss; r1.tm = sm;
    r1((long long)threadIdx.x);
    __syncthreads();
    if (threadIdx.x == 0) {
        ReduceParts2 r2{ss, sm, out};
        r2(0);
    }
}
#endif
static void reduce_parts(stream_t s, const double* part_sum, const float* part_max, int nparts, double* scratch_sum, float* scratch_max, double* out) {
#ifndef MVD_HOST_EMU
    (void)scratch_sum; (void)scratch_max;
    reduce_parts_kernel<<<1, 256, 0, s>>>(ReduceParts1{part_sum, part_max, nparts, nullptr, nullptr}, out);
    MVD_CUDA_CHECK(cudaGetLastError());
#else
    pfor(256, ReduceParts1{part_sum, part_max, nparts, scratch_sum, scratch_max}, s);
    pfor(1, ReduceParts2{scratch_sum, scratch_max, out}, s);
#endif
}

// ------------------------------------------------------------------------------------------------
// Convolver
// ------------------------------------------------------------------------------------------------
Convolver::Convolver(const Geometry& g, const Reach r1[3], const Reach r2[3], int xmode, int max_len, stream_t s, Tables* tables,
                     bool two_exchanges)
    : g_(g), xmode_(xmode), stream_(s), tables_(tables) {
    if (const char* e = std::getenv("MVD_PREFETCH_DIST")) pf_x_ = pf_y_ = pf_z_ = std::atoi(e);
    if (const char* e = std::getenv("MVD_PF_X")) pf_x_ = std::atoi(e);
    if (const char* e = std::getenv("MVD_PF_Y")) pf_y_ = std::atoi(e);
    if (const char* e = std::getenv("MVD_PF_Z")) pf_z_ = std::atoi(e);
    AxisTiling ax[3];
    if (xmode == 1) {
        for (int d = 0; d < 3; ++d) {
            if (!find_len_ops(g.gdim[d])) throw Error("circular convolution: dimension " + std::to_string(g.gdim[d]) + " is not a supported FFT length");
            ax[d].T = g.gdim[d];
            ax[d].tiles.push_back(AxisTile{0, 0, g.gdim[d]});
        }
        M_ = ax[0].T;
    } else {
        for (int d = 0; d < 3; ++d) {
            const bool shard = g.own_lo[d] != 0 || g.own_hi[d] != g.gdim[d];
            ax[d] = plan_axis(g.gdim[d], g.own_lo[d], g.own_hi[d], r1[d], r2[d], d == 0, max_len, two_exchanges && shard);
        }
        M_ = ax[0].T / 2;
    }
    for (int d = 0; d < 3; ++d) T_[d] = ax[d].T;
    two_z_ = xmode != 1 && two_exchanges && (g.own_lo[2] != 0 || g.own_hi[2] != g.gdim[2]);
    r2z_ = r2[2];
    r2y_ = r2[1];
    for (int d = 0; d < 3; ++d) { shard_lo_[d] = g.own_lo[d] != 0; shard_hi_[d] = g.own_hi[d] != g.gdim[d]; }
    ox_ = find_len_ops(M_);
    oy_ = find_len_ops(T_[1]);
    oz_ = find_len_ops(T_[2]);
    if (!ox_ || !oy_ || !oz_) throw Error("internal: planned length without kernels");
    px_ = (M_ + 15) / 16 * 16;                            // 128-byte rows: every 16-column segment is one cache line
    xblocks_ = (T_[1] * T_[2] + ox_->XL - 1) / ox_->XL + T_[2];      // partial-statistics slots per tile (chunked launches included)
    {
        double mb = 0.0;      // MVD_CHUNK_MB: plane-chunk size (MB) of the x/y chains; measured slower than whole-tile launches on B200 -> off
        if (const char* e = std::getenv("MVD_CHUNK_MB")) mb = std::atof(e);
        const double plane_mb = (double)px_ * T_[1] * sizeof(cpx) / 1.0e6;
        chunk_planes_ = mb > 0 ? std::max(1, (int)(mb / plane_mb)) : 0;
        if (chunk_planes_ >= T_[2]) chunk_planes_ = 0;
    }
    for (const AxisTile& tz : ax[2].tiles)
        for (const AxisTile& ty : ax[1].tiles)
            for (const AxisTile& tx : ax[0].tiles) {
                TileGeom t;
                t.org[0] = tx.org; t.lo[0] = tx.lo; t.hi[0] = tx.hi;
                t.org[1] = ty.org; t.lo[1] = ty.lo; t.hi[1] = ty.hi;
                t.org[2] = tz.org; t.lo[2] = tz.lo; t.hi[2] = tz.hi;
                tiles_.push_back(t);
            }
    work_ = (cpx*)dev::alloc(sizeof(cpx) * tile_elems());
    dev::zero(work_, sizeof(cpx) * tile_elems(), stream_);   // pitch padding columns stay finite
}
Convolver::~Convolver() {
    dev::free_(work_);
    dev::free_(kpad_);
    dev::free_(kdev_);
#ifndef MVD_HOST_EMU
    for (void* e : prof_events_) cudaEventDestroy((cudaEvent_t)e);
#endif
}
double Convolver::fft_volume_ratio() const {
    double useful = 0;
    for (const TileGeom& t : tiles_) useful += (double)(t.hi[0] - t.lo[0]) * (t.hi[1] - t.lo[1]) * (t.hi[2] - t.lo[2]);
    return (double)tiles_.size() * T_[0] * T_[1] * T_[2] / std::max(useful, 1.0);
}

XArgs Convolver::base_xargs(const TileGeom& t) const {
    XArgs a;
    std::memset(&a, 0, sizeof(a));
    a.cdata = work_;
    a.px = px_;
    a.nlines = T_[1] * T_[2];
    a.ty = T_[1];
    a.tw = tables_->xtw(M_);
    a.twist = xmode_ == 0 ? tables_->twist(M_) : nullptr;
    a.xmode = xmode_;
    for (int d = 0; d < 3; ++d) {
        a.vol[d] = g_.vol[d]; a.gdim[d] = g_.gdim[d]; a.goff[d] = g_.goff[d];
        a.org[d] = t.org[d]; a.vlo[d] = t.lo[d]; a.vhi[d] = t.hi[d];
    }
    a.ext = EXT_MIRROR;
    a.ext_value = 0.f;
    a.min_value = 1e-4f;
    a.max_intensity = 1.f;
    a.pf_dist = pf_x_;
    a.nblocks = xblocks_;
    // one reflection suffices when the tile never reaches further than gdim-1 beyond either face of the volume
    a.xsimple = (t.org[0] >= -(g_.gdim[0] - 1) && t.org[0] + T_[0] - 1 <= 2 * g_.gdim[0] - 2) ? 1 : 0;
    return a;
}

void Convolver::xpass(int kind, XArgs a, int z0, int z1) {
    if (z1 < 0) z1 = T_[2];
    a.line0 = z0 * T_[1];
    a.line_end = z1 * T_[1];
    // MVD_CHECK_RECTS=1 (tests): the rectangle list of a filtered launch must enumerate exactly the lines the box filter selects --
    // the two forms are consumed by different kernel shapes and have to describe the same set
    static const bool check_rects = [] { const char* e = std::getenv("MVD_CHECK_RECTS"); return e && std::atoi(e) != 0; }();
    if (check_rects && a.nrect > 0) {
        a.ty = T_[1];
        std::vector<unsigned char> hit((size_t)T_[1] * T_[2], 0);
        for (int r = 0; r < a.nrect; ++r) {
            if (a.rect_start[r + 1] - a.rect_start[r] != (a.rect[r][1] - a.rect[r][0]) * (a.rect[r][3] - a.rect[r][2])) throw Error("internal: selection rectangle count");
            for (int z = a.rect[r][2]; z < a.rect[r][3]; ++z)
                for (int y = a.rect[r][0]; y < a.rect[r][1]; ++y) {
                    if (y < 0 || y >= T_[1] || z < 0 || z >= T_[2] || hit[(size_t)z * T_[1] + y]) throw Error("internal: selection rectangles overlap or leave the tile");
                    hit[(size_t)z * T_[1] + y] = 1;
                }
        }
        for (int l = 0; l < T_[1] * T_[2]; ++l) {
            const bool want = l >= a.line0 && x_line_selected(a, l);
            if (want != (hit[(size_t)l] != 0)) throw Error("internal: selection rectangles and line filter disagree at line " + std::to_string(l));
        }
    }
    const int nb = (a.line_end - a.line0 + ox_->XL - 1) / ox_->XL;
    a.nblocks = nb;
    if (a.part_sum) { a.part_sum += (size_t)(a.line0 / ox_->XL) + (size_t)z0; a.part_max += (size_t)(a.line0 / ox_->XL) + (size_t)z0; }
    ox_->launch_x(kind, a, nb, stream_);
}

void Convolver::col(int axis, int mode, const cpx* khat, int z0, int z1) {
    ColArgs c;
    c.data = work_;
    c.khat = khat;
    c.nx = M_;
    const LenOps* o = axis == 1 ? oy_ : oz_;
    c.tw = tables_->tw(o->N);
    int gy;
    if (axis == 1) {                                     // y pass: one CTA row per z plane; plane chunks [z0, z1)
        if (z1 < 0) z1 = T_[2];
        c.stride_n = px_; c.stride_b = (long long)px_ * T_[1]; gy = z1 - z0;
        c.data = work_ + (size_t)z0 * c.stride_b;
    }
    else { c.stride_n = (long long)px_ * T_[1]; c.stride_b = px_; gy = T_[1]; }
    c.gx = (M_ + o->W - 1) / o->W;
    c.gy = gy;
    c.pf_dist = axis == 1 ? pf_y_ : pf_z_;
    o->launch_col(mode, c, c.gx, gy, stream_);
}

cpx* Convolver::build_khat(const float* kernel_host, const int kd[3]) {
    cpx* khat = (cpx*)dev::alloc(sizeof(cpx) * tile_elems());
    try { build_khat_into(khat, kernel_host, kd); } catch (...) { dev::free_(khat); throw; }
    return khat;
}

// No cudaFree and no synchronisation in here: cudaFree waits for the whole device, i.e. also for the view uploads that run on the copy
// stream while the spectra are built (DeconViews with async upload).  The staging buffers live as long as the plan.
void Convolver::build_khat_into(cpx* khat, const float* kernel_host, const int kd[3]) {
    for (int d = 0; d < 3; ++d)
        if (kd[d] > T_[d] || kd[d] < 1) throw Error("kernel larger than the FFT tile");
    const size_t nk = (size_t)kd[0] * kd[1] * kd[2];
    const size_t nt = (size_t)T_[0] * T_[1] * T_[2];
    if (!kpad_) kpad_ = (float*)dev::alloc(sizeof(float) * nt);
    if (nk > kdev_cap_) {
        float* bigger = (float*)dev::alloc(sizeof(float) * nk);
        if (kdev_) { dev::sync(stream_); dev::free_(kdev_); }
        kdev_ = bigger; kdev_cap_ = nk;
    }
    dev::h2d(kdev_, kernel_host, sizeof(float) * nk, stream_);          // pageable source: the call returns once the data is staged
    dev::zero(kpad_, sizeof(float) * nt, stream_);
    const double nfft = (double)M_ * (double)T_[1] * (double)T_[2];
    PlaceKernel pk{kdev_, kpad_, kd[0], kd[1], kd[2], T_[0], T_[1], T_[2], (float)(1.0 / nfft), xmode_ == 0 ? 1 : 0};
    pfor((long long)nk, pk, stream_);
    // forward transform of the padded kernel through the very same passes (scrambled order matches by construction)
    TileGeom t;
    for (int d = 0; d < 3; ++d) { t.org[d] = 0; t.lo[d] = 0; t.hi[d] = T_[d]; }
    XArgs a = base_xargs(t);
    for (int d = 0; d < 3; ++d) { a.vol[d] = T_[d]; a.gdim[d] = T_[d]; a.goff[d] = 0; }
    a.src = kpad_;
    a.ext = EXT_ZERO;
    xpass(X_FWD, a);
    col(1, COL_FWD, nullptr);
    col(2, COL_FWD, nullptr);
    dev::d2d(khat, work_, sizeof(cpx) * tile_elems(), stream_);
}

void Convolver::conv(const float* src, float* dst, const cpx* khat, int ext, float ext_value) {
    for (const TileGeom& t : tiles_) {
        XArgs a = base_xargs(t);
        a.src = src;
        a.ext = ext;
        a.ext_value = ext_value;
        xpass(X_FWD, a);
        col(1, COL_FWD, nullptr);
        col(2, COL_CONV, khat);
        col(1, COL_INV, nullptr);
        a.dst = dst;
        xpass(X_INV, a);
    }
}

// selection rectangles {y0, y1, z0, z1} of a filtered x launch (XArgs::rect): kernels that deal single lines enumerate them
static void rects_clear(XArgs& a) { a.nrect = 0; a.rect_start[0] = 0; }
static void rects_add(XArgs& a, int y0, int y1, int z0, int z1) {
    if (y1 <= y0 || z1 <= z0 || a.nrect >= 8) return;
    int* r = a.rect[a.nrect];
    r[0] = y0; r[1] = y1; r[2] = z0; r[3] = z1;
    a.rect_start[a.nrect + 1] = a.rect_start[a.nrect] + (y1 - y0) * (z1 - z0);
    ++a.nrect;
}
// outer box minus inner box (inner clipped to outer) as up to four rectangles: the z slabs below / above, the y strips beside
static void rects_add_difference(XArgs& a, const int o[4], const int in[4]) {
    const int iy0 = std::max(o[0], in[0]), iy1 = std::min(o[1], in[1]), iz0 = std::max(o[2], in[2]), iz1 = std::min(o[3], in[3]);
    if (iy1 <= iy0 || iz1 <= iz0) { rects_add(a, o[0], o[1], o[2], o[3]); return; }
    rects_add(a, o[0], o[1], o[2], iz0);
    rects_add(a, o[0], o[1], iz1, o[3]);
    rects_add(a, o[0], iy0, iz0, iz1);
    rects_add(a, iy1, o[1], iz0, iz1);
}

void Convolver::view_update(const float* psi_in, float* psi_out, const float* img, const float* weight, const cpx* k1hat,
                            const cpx* k2hat, float lambda, float min_value, float max_intensity, double* part_sum, float* part_max,
                            const unsigned char* skip, const std::function<void()>* psi_join) {
    if (xmode_ != 0) throw Error("internal: view updates need the real-packed x mode");
    int ti = 0;
    const int cp = chunk_planes_ > 0 ? chunk_planes_ : T_[2];
    for (const TileGeom& t : tiles_) {
        if (skip && skip[ti]) {
            // no content: psi keeps its values inside the tile's responsibility box (full x rows: tiles that split x are copied more than once)
            copy_region(stream_, psi_in, g_.vol[1], t.lo[1] - g_.goff[1], t.lo[2] - g_.goff[2], psi_out, g_.vol[1], t.lo[1] - g_.goff[1],
                        t.lo[2] - g_.goff[2], g_.vol[0], t.hi[1] - t.lo[1], t.hi[2] - t.lo[2]);
            clear_parts(stream_, part_sum + (size_t)ti * xblocks_, part_max + (size_t)ti * xblocks_, xblocks_);
            ++ti;
            continue;
        }
        XArgs a = base_xargs(t);
        // x/y pass chains run plane chunk by plane chunk so that the hand-over P1->P2, P4->P5->P6, P8->P9 stays in the 126 MB L2;
        // the z passes (P3, P7) need every plane and run over the whole tile.
        // own box of this rank in tile coordinates
        const int oy0 = std::max(0, g_.own_lo[1] - t.org[1]), oy1 = std::min(T_[1], g_.own_hi[1] - t.org[1]);
        const int oz0 = std::max(0, g_.own_lo[2] - t.org[2]), oz1 = std::min(T_[2], g_.own_hi[2] - t.org[2]);
        a.src = psi_in;                                   // P1: mirror-single outside the volume (MultiViewDeconvolutionSeq.java:115)
        a.ext = EXT_MIRROR;
        // MVD_SPLIT_P1=1: measured on c3 the two launches cost more (0.13 ms per view update at N = 2, 0.06 ms at N = 8) than the part of
        // the psi exchange they hide -- most of it already travels behind the statistics kernel -- so the split is opt-in
        static const bool split_p1 = [] { const char* e = std::getenv("MVD_SPLIT_P1"); return e && std::atoi(e) != 0; }();
        if (psi_join && *psi_join && chunk_planes_ == 0 && split_p1) {
            // the halo exchange of psi is still travelling: the lines of the own box do not read halo data (a reflected row at a
            // volume face is an own row) and are transformed first; everything else follows once the halos are in place
            const int own[4] = {oy0, oy1, oz0, oz1}, whole[4] = {0, T_[1], 0, T_[2]};
            a.fin[0] = oy0; a.fin[1] = oy1; a.fin[2] = oz0; a.fin[3] = oz1;
            rects_clear(a); rects_add(a, oy0, oy1, oz0, oz1);
            mark(0); xpass(X_FWD, a, oz0, oz1);
            mark(11);
            (*psi_join)();
            psi_join = nullptr;
            a.fin[0] = a.fin[1] = 0;
            a.fout[0] = oy0; a.fout[1] = oy1; a.fout[2] = oz0; a.fout[3] = oz1;
            rects_clear(a); rects_add_difference(a, whole, own);
            mark(0); xpass(X_FWD, a);
            a.fout[0] = a.fout[1] = 0;
            rects_clear(a);
            mark(1); col(1, COL_FWD, nullptr);            // P2
        } else {
            if (psi_join && *psi_join) { (*psi_join)(); psi_join = nullptr; }
            for (int z0 = 0; z0 < T_[2]; z0 += cp) {
                const int z1 = std::min(T_[2], z0 + cp);
                mark(0); xpass(X_FWD, a, z0, z1);
                mark(1); col(1, COL_FWD, nullptr, z0, z1);    // P2
            }
        }
        mark(2); col(2, COL_CONV, k1hat);                 // P3
        // Planes the quotient is computed on.  Exchange scheme 1 on a z-sharded box: the quotient of the halo planes on an interior
        // side arrives (as its x-spectrum) from the z neighbour, so P4 / P5 skip them; planes beyond the reach of kernel2 are not
        // read by anything that ends up in the responsibility box and are cleared (sane values for the transforms that follow).
        // Sides that are volume faces keep all their planes: outside the volume the quotient is 1, and P5 writes exactly that.
        int q0 = 0, q1 = T_[2];
        if (two_z_ && mid_exchange_ && chunk_planes_ == 0) {
            const size_t plane_bytes = sizeof(cpx) * (size_t)px_ * T_[1];
            if (g_.own_lo[2] != 0) {
                q0 = g_.own_lo[2] - t.org[2];
                const int keep = std::min(q0, r2z_.lo);
                if (q0 - keep > 0) dev::zero(work_, plane_bytes * (size_t)(q0 - keep), stream_);
            }
            if (g_.own_hi[2] != g_.gdim[2]) {
                q1 = g_.own_hi[2] - t.org[2];
                const int first = std::min(T_[2], q1 + r2z_.hi);
                if (T_[2] - first > 0) dev::zero(work_ + (size_t)px_ * T_[1] * first, plane_bytes * (size_t)(T_[2] - first), stream_);
            }
        }
        static const bool split_p5 = [] { const char* e = std::getenv("MVD_SPLIT_P5"); return !(e && std::atoi(e) == 0); }();   // MVD_SPLIT_P5=0: A/B
        if (mid_exchange_ && mid_join_ && chunk_planes_ == 0 && split_p5) {
            // Boundary-first quotient pass.  This rank computes the quotient on its own box, extended to the tile edge where the box
            // ends at a volume face (outside the volume the quotient is 1, and P5 writes exactly that); the neighbours deliver the
            // rest within the reach of kernel2, what lies beyond is cleared.  The lines the neighbours are waiting for -- own lines
            // within that reach of an interior side -- run first, then the exchange starts and travels while the deep interior
            // follows.
            const int ey0 = shard_lo_[1] ? oy0 : 0, ey1 = shard_hi_[1] ? oy1 : T_[1];
            mark(3); col(1, COL_INV, nullptr, q0, q1);    // P4 (own planes; face sides keep theirs)
            mark(11);
            const size_t row_bytes = sizeof(cpx) * (size_t)px_, plane_bytes = row_bytes * (size_t)T_[1];
            if (shard_lo_[1] && ey0 - r2y_.lo > 0)
                dev::zero2d(work_ + (size_t)px_ * T_[1] * q0, plane_bytes, row_bytes * (size_t)(ey0 - r2y_.lo), (size_t)(q1 - q0), stream_);
            if (shard_hi_[1] && ey1 + r2y_.hi < T_[1])
                dev::zero2d(work_ + (size_t)px_ * ((size_t)T_[1] * q0 + (size_t)(ey1 + r2y_.hi)), plane_bytes,
                            row_bytes * (size_t)(T_[1] - ey1 - r2y_.hi), (size_t)(q1 - q0), stream_);
            // deep interior: further than the neighbours' reach from every interior side
            const int dy0 = ey0 + (shard_lo_[1] ? r2y_.hi : 0), dy1 = ey1 - (shard_hi_[1] ? r2y_.lo : 0);
            const int dz0 = q0 + (shard_lo_[2] ? r2z_.hi : 0), dz1 = q1 - (shard_hi_[2] ? r2z_.lo : 0);
            const bool deep = dy1 > dy0 && dz1 > dz0;
            a.src = img;                                  // P5: quotient, 1 where there is no image data
            a.fin[0] = ey0; a.fin[1] = ey1; a.fin[2] = q0; a.fin[3] = q1;
            a.fkeep = 1;      // + all rows of the planes outside the volume in z (face margins): their quotient is 1, nobody delivers them
            if (deep) { a.fout[0] = dy0; a.fout[1] = dy1; a.fout[2] = dz0; a.fout[3] = dz1; }
            {
                // the same selection as rectangles: (own box, extended at faces) minus the deep interior, plus the interior-side halo rows
                // of the planes that lie outside the volume in z
                const int ext[4] = {ey0, ey1, q0, q1}, dp[4] = {dy0, dy1, dz0, dz1};
                rects_clear(a);
                if (deep) rects_add_difference(a, ext, dp); else rects_add(a, ey0, ey1, q0, q1);
                const int vz0 = std::min(q1, std::max(q0, -t.org[2])), vz1 = std::max(q0, std::min(q1, g_.gdim[2] - t.org[2]));
                rects_add(a, 0, ey0, q0, vz0); rects_add(a, ey1, T_[1], q0, vz0);
                rects_add(a, 0, ey0, vz1, q1); rects_add(a, ey1, T_[1], vz1, q1);
            }
            mark(4); xpass(X_RATIO, a, q0, q1);
            mark(9);
            a.fkeep = 0;
            rects_clear(a);
            mid_exchange_(work_, t);
            if (deep) {
                a.fout[0] = a.fout[1] = 0;
                a.fin[0] = dy0; a.fin[1] = dy1; a.fin[2] = dz0; a.fin[3] = dz1;
                rects_add(a, dy0, dy1, dz0, dz1);
                mark(4); xpass(X_RATIO, a, dz0, dz1);
                mark(9);
                rects_clear(a);
            }
            a.fin[0] = a.fin[1] = a.fout[0] = a.fout[1] = 0;
            mid_join_();
            mark(5); col(1, COL_FWD, nullptr);            // P6
        } else {
            for (int z0 = q0; z0 < q1; z0 += cp) {
                const int z1 = std::min(q1, z0 + cp);
                mark(3); col(1, COL_INV, nullptr, z0, z1);    // P4
                a.src = img;                                  // P5: quotient, 1 where there is no image data
                mark(4); xpass(X_RATIO, a, z0, z1);
                if (!mid_exchange_) { mark(5); col(1, COL_FWD, nullptr, z0, z1); }   // P6
            }
            if (mid_exchange_) {                              // scheme B: the neighbours' quotient rows / planes arrive as x-spectra
                mark(9);
                mid_exchange_(work_, t);
                for (int z0 = 0; z0 < T_[2]; z0 += cp) { mark(5); col(1, COL_FWD, nullptr, z0, std::min(T_[2], z0 + cp)); }
            }
        }
        mark(6); col(2, COL_CONV, k2hat);                 // P7
        a.src = psi_in;                                   // P9
        a.weight = weight;
        a.dst = psi_out;
        a.lambda = lambda;
        a.min_value = min_value;
        a.max_intensity = max_intensity;
        a.part_sum = part_sum + (size_t)ti * xblocks_;
        a.part_max = part_max + (size_t)ti * xblocks_;
        // the update only touches the planes of the responsibility box: the inverse y pass before it is not needed anywhere else
        const int u0 = chunk_planes_ == 0 ? std::max(0, t.lo[2] - t.org[2]) : 0, u1 = chunk_planes_ == 0 ? std::min(T_[2], t.hi[2] - t.org[2]) : T_[2];
        for (int z0 = u0; z0 < u1; z0 += cp) {
            const int z1 = std::min(u1, z0 + cp);
            mark(7); col(1, COL_INV, nullptr, z0, z1);    // P8
            mark(8); xpass(X_UPDATE, a, z0, z1);
        }
        mark(10);
        ++ti;
    }
}

void Convolver::integral(const float* psi_in, const float* img, const cpx* k1hat, const cpx* k2hat, float* integral_out) {
    if (xmode_ != 0) throw Error("internal: the quotient pass needs the real-packed x mode");
    for (const TileGeom& t : tiles_) {
        XArgs a = base_xargs(t);
        a.src = psi_in;
        a.ext = EXT_MIRROR;
        xpass(X_FWD, a);
        col(1, COL_FWD, nullptr); col(2, COL_CONV, k1hat); col(1, COL_INV, nullptr);
        a.src = img;
        xpass(X_RATIO, a);
        col(1, COL_FWD, nullptr); col(2, COL_CONV, k2hat); col(1, COL_INV, nullptr);
        a.dst = integral_out;
        xpass(X_INV, a);
    }
}

// ------------------------------------------------------------------------------------------------
// kernel helpers (host; kernels are a few thousand voxels)
// ------------------------------------------------------------------------------------------------
std::vector<float> mirror_kernel(const std::vector<float>& k, const int kd[3]) {
    // Mirror.mirror (M/process/deconvolution/util/Mirror.java:96-108) applied to every axis: positions p <= dim/2 are swapped
    // with dim-1-p, so for an even size the two middle samples are swapped twice and stay put ("Quirk C").
    std::vector<float> cur = k;
    for (int axis = 0; axis < 3; ++axis) {
        std::vector<float> out(cur.size());
        const int n = kd[axis];
        for (int z = 0; z < kd[2]; ++z)
            for (int y = 0; y < kd[1]; ++y)
                for (int x = 0; x < kd[0]; ++x) {
                    int c[3] = {x, y, z};
                    int p = c[axis];
                    int q = n - 1 - p;
                    if (n % 2 == 0 && (p == n / 2 || p == n / 2 - 1)) q = p;
                    int s[3] = {x, y, z};
                    s[axis] = q;
                    out[((size_t)z * kd[1] + y) * kd[0] + x] = cur[((size_t)s[2] * kd[1] + s[1]) * kd[0] + s[0]];
                }
        cur.swap(out);
    }
    return cur;
}
double sum_kernel(const std::vector<float>& k) {
    // RealSum-like compensated summation (AdjustInput.sumImg, AdjustInput.java:65-122; exact-sum policy, no double count)
    double s = 0.0, c = 0.0;
    for (float v : k) { double y = (double)v - c; double t = s + y; c = (t - s) - y; s = t; }
    return s;
}
void norm_to_sum1(std::vector<float>& k, int quirk_threads) {   // AdjustInput.normToSum1 (AdjustInput.java:52-58)
    double s = sum_kernel(k);
    if (quirk_threads > 0 && !k.empty()) {
        // "Quirk A": sums[0] is added once before the loop over all portion sums (AdjustInput.java:115-119)
        const long long size = (long long)k.size();
        const long long T = std::max(4, quirk_threads);                                  // Threads.numThreads()
        long long np = size <= T ? size : std::max(T, size / (64LL * 64LL * 64LL));      // FusionTools.divideIntoPortions
        long long chunk = size / np;
        while (chunk == 0) { --np; chunk = size / np; }
        std::vector<float> first(k.begin(), k.begin() + chunk);
        s += sum_kernel(first);
    }
    for (float& v : k) v = (float)((double)v / s);
}

// ------------------------------------------------------------------------------------------------
// Engine
// ------------------------------------------------------------------------------------------------
Engine::Engine(const Config& c) : cfg_(c) {
    if (c.num_views < 1) throw Error("need at least one view");
    for (int d = 0; d < 3; ++d) {
        const Geometry& g = c.geom;
        if (g.vol[d] < 1 || g.gdim[d] < 1) throw Error("empty volume");
        if (g.own_lo[d] < 0 || g.own_hi[d] > g.gdim[d] || g.own_lo[d] >= g.own_hi[d]) throw Error("bad responsibility box");
        if (g.own_lo[d] < g.goff[d] || g.own_hi[d] > g.goff[d] + g.vol[d]) throw Error("responsibility box outside the local array");
    }
    dev::set_device(c.device);
    stream_ = dev::stream_create();
    tables_.reset(new Tables(stream_));
    views_.resize(c.num_views);
    const size_t bytes = sizeof(float) * local_voxels();
    psi_[0] = (float*)dev::alloc(bytes);
    psi_[1] = (float*)dev::alloc(bytes);
    dev::zero(psi_[0], bytes, stream_);
    dev::zero(psi_[1], bytes, stream_);
}

Engine::~Engine() {
    try { dev::set_device(cfg_.device); dev::sync(stream_); } catch (...) {}
    if (copy_stream_) { try { dev::sync(copy_stream_); } catch (...) {} }
    for (View& v : views_) {
        dev::free_(v.img_owned); dev::free_(v.weight_owned); dev::free_(v.k1hat); dev::free_(v.k2hat);
        dev::event_destroy(v.ready);
    }
    small_conv_.reset();
    if (xstream_) { try { dev::sync(xstream_); } catch (...) {} }
    dev::stream_destroy(xstream_);
    dev::event_destroy(ev_compute_); dev::event_destroy(ev_psi_); dev::event_destroy(ev_mid_);
    dev::stream_destroy(copy_stream_);
    dev::free_(psi_[0]); dev::free_(psi_[1]);
    dev::free_(part_sum_); dev::free_(part_max_); dev::free_(stats_dev_);
    comm_.reset();
    for (float* p : integral_) dev::free_(p);
    dev::free_(lut_dev_); dev::free_(acc_dev_); dev::free_(max_dev_); dev::free_(flag_dev_);
    dev::free_(small_buf_); dev::free_(small_khat_);
    conv_.reset();
    tables_.reset();
    dev::stream_destroy(stream_);
}

void Engine::set_view_host(int v, const float* img, const float* weight) {
    if (v < 0 || v >= cfg_.num_views) throw Error("view index out of range");
    dev::set_device(cfg_.device);
    View& vw = views_[v];
    wait_upload(vw);
    const size_t bytes = sizeof(float) * local_voxels();
    if (!vw.img_owned) vw.img_owned = (float*)dev::alloc(bytes);
    if (!vw.weight_owned) vw.weight_owned = (float*)dev::alloc(bytes);
    dev::h2d(vw.img_owned, img, bytes, stream_);
    if (weight) dev::h2d(vw.weight_owned, weight, bytes, stream_);       // weight == nullptr: generated on the device later
    else dev::zero(vw.weight_owned, bytes, stream_);
    dev::sync(stream_);       // the synchronous variant: the caller's buffers (possibly page-locked) are free again when this returns
    vw.pending = false;
    vw.img = vw.img_owned;
    vw.weight = vw.weight_owned;
    weights_changed();
}
void Engine::set_view_host_async(int v, const float* img, const float* weight) {
    if (v < 0 || v >= cfg_.num_views) throw Error("view index out of range");
    dev::set_device(cfg_.device);
    View& vw = views_[v];
    const size_t bytes = sizeof(float) * local_voxels();
    if (!vw.img_owned) vw.img_owned = (float*)dev::alloc(bytes);
    if (!vw.weight_owned) vw.weight_owned = (float*)dev::alloc(bytes);
    if (!copy_stream_) copy_stream_ = dev::stream_create();
    if (!vw.ready) vw.ready = dev::event_create();
    dev::h2d(vw.img_owned, img, bytes, copy_stream_);
    if (weight) dev::h2d(vw.weight_owned, weight, bytes, copy_stream_);
    else dev::zero(vw.weight_owned, bytes, copy_stream_);
    dev::event_record(vw.ready, copy_stream_);
    vw.pending = true;
    vw.img = vw.img_owned;
    vw.weight = vw.weight_owned;
    weights_changed();
}
void Engine::set_view_device(int v, const float* img, const float* weight) {
    if (v < 0 || v >= cfg_.num_views) throw Error("view index out of range");
    View& vw = views_[v];
    vw.img = img;
    if (weight) { vw.weight = weight; weights_changed(); return; }
    // weight == nullptr: the weight mask will be generated on the device (mvd_make_blending_weights + mvd_normalize_weights)
    dev::set_device(cfg_.device);
    if (!vw.weight_owned) vw.weight_owned = (float*)dev::alloc(sizeof(float) * local_voxels());
    dev::zero(vw.weight_owned, sizeof(float) * local_voxels(), stream_);
    vw.weight = vw.weight_owned;
    weights_changed();
}
void Engine::set_psf(int v, const float* psf, const int kd[3]) {
    if (v < 0 || v >= cfg_.num_views) throw Error("view index out of range");
    View& vw = views_[v];
    const size_t n = (size_t)kd[0] * kd[1] * kd[2];
    vw.psf.assign(psf, psf + n);
    for (int d = 0; d < 3; ++d) vw.psf_dims[d] = kd[d];
    vw.k1.clear(); vw.k2.clear();
    inited_ = false;
}
void Engine::set_kernels(int v, const float* k1, const int k1d[3], const float* k2, const int k2d[3]) {
    if (v < 0 || v >= cfg_.num_views) throw Error("view index out of range");
    View& vw = views_[v];
    vw.k1.assign(k1, k1 + (size_t)k1d[0] * k1d[1] * k1d[2]);
    vw.k2.assign(k2, k2 + (size_t)k2d[0] * k2d[1] * k2d[2]);
    for (int d = 0; d < 3; ++d) { vw.k1d[d] = k1d[d]; vw.k2d[d] = k2d[d]; }
    vw.psf.clear();
    inited_ = false;
}

std::vector<float> Engine::conv_same(const std::vector<float>& in, const int d[3], const std::vector<float>& k, const int kd[3]) {
    // zero-extended "same"-size convolution of two PSF-sized volumes (DeconViewPSF.java:152-178,215-225); the tile plan is cached
    // because every call of one derivation has the same extents
    std::vector<float> out(in.size());
    bool same = small_conv_ != nullptr;
    for (int a = 0; a < 3; ++a) same = same && small_dims_[a] == d[a] && small_kd_[a] == kd[a];
    if (!same) {
        Geometry g;
        Reach r1[3], r2[3];
        for (int a = 0; a < 3; ++a) {
            g.gdim[a] = g.vol[a] = d[a]; g.goff[a] = 0; g.own_lo[a] = 0; g.own_hi[a] = d[a];
            r1[a] = reach_of(kd[a]); r2[a] = Reach{0, 0};
            small_dims_[a] = d[a]; small_kd_[a] = kd[a];
        }
        small_conv_.reset(new Convolver(g, r1, r2, 0, cfg_.max_len, stream_, tables_.get()));
    }
    const size_t n = in.size();
    if (2 * n > small_buf_cap_ || small_conv_->tile_elems() > small_khat_cap_) {       // persistent scratch: no cudaFree per call (see build_khat_into)
        dev::sync(stream_);
        dev::free_(small_buf_); dev::free_(small_khat_);
        small_buf_ = (float*)dev::alloc(sizeof(float) * n * 2); small_buf_cap_ = 2 * n;
        small_khat_ = (cpx*)dev::alloc(sizeof(cpx) * small_conv_->tile_elems()); small_khat_cap_ = small_conv_->tile_elems();
    }
    float* s = small_buf_;
    dev::h2d(s, in.data(), sizeof(float) * n, stream_);
    small_conv_->build_khat_into(small_khat_, k.data(), kd);
    small_conv_->conv(s, s + n, small_khat_, EXT_ZERO, 0.f);
    dev::d2h(out.data(), s + n, sizeof(float) * n, stream_);
    dev::sync(stream_);
    return out;
}

// DeconViewPSF.init for every view in list order (DeconViewPSF.java:119-254, DeconViews.java:69-70)
void Engine::derive_kernels() {
    const int V = cfg_.num_views;
    bool any_psf = false;
    for (View& vw : views_) any_psf = any_psf || !vw.psf.empty();
    if (!any_psf) return;
    for (View& vw : views_) {
        if (vw.psf.empty()) throw Error("either all views get a PSF (set_psf) or all get explicit kernels (set_kernels)");
        vw.k1 = vw.psf;                       // not yet normalised -- see Quirk B below
        for (int d = 0; d < 3; ++d) { vw.k1d[d] = vw.psf_dims[d]; vw.k2d[d] = vw.psf_dims[d]; }
    }
    for (int v = 0; v < V; ++v) {
        View& me = views_[v];
        norm_to_sum1(me.k1, cfg_.norm_quirk_threads);                                            // :125
        if (V == 1 || cfg_.psf_type == INDEPENDENT) {                                            // :127-131
            me.k2 = mirror_kernel(me.k1, me.k1d);
        } else if (cfg_.psf_type == EFFICIENT_BAYESIAN) {                                        // :132-195
            std::vector<float> tmp = mirror_kernel(me.k1, me.k1d);
            for (int w = 0; w < V; ++w) {
                if (w == v) continue;
                // other views' kernel1 is normalised only if their init already ran (w < v): "Quirk B", harmless
                View& ot = views_[w];
                std::vector<float> out = conv_same(mirror_kernel(me.k1, me.k1d), me.k1d, ot.k1, ot.k1d);
                out = conv_same(out, me.k1d, mirror_kernel(ot.k1, ot.k1d), ot.k1d);
                for (size_t i = 0; i < tmp.size(); ++i) tmp[i] = out[i] * tmp[i];
            }
            norm_to_sum1(tmp, cfg_.norm_quirk_threads);
            me.k2 = tmp;
        } else if (cfg_.psf_type == OPTIMIZATION_I) {                                            // :196-242
            std::vector<float> tmp = me.k1;
            for (int w = 0; w < V; ++w) {
                if (w == v) continue;
                View& ot = views_[w];
                std::vector<float> out = conv_same(me.k1, me.k1d, mirror_kernel(ot.k1, ot.k1d), ot.k1d);
                for (size_t i = 0; i < tmp.size(); ++i) tmp[i] = out[i] * tmp[i];
            }
            norm_to_sum1(tmp, cfg_.norm_quirk_threads);
            me.k2 = mirror_kernel(tmp, me.k1d);
        } else {                                                                                 // OPTIMIZATION_II :243-253
            std::vector<float> e = me.k1;
            for (size_t i = 0; i < e.size(); ++i) {
                float r = me.k1[i];
                for (int p = 1; p < V; ++p) r *= me.k1[i];       // pow by repeated float multiply (:276-284)
                e[i] = r;
            }
            norm_to_sum1(e, cfg_.norm_quirk_threads);
            me.k2 = mirror_kernel(e, me.k1d);
        }
    }
}

void Engine::init_views() {
    dev::set_device(cfg_.device);
    derive_kernels();
    Reach r1[3] = {{0, 0}, {0, 0}, {0, 0}}, r2[3] = {{0, 0}, {0, 0}, {0, 0}};
    for (View& vw : views_) {
        if (vw.k1.empty() || vw.k2.empty()) throw Error("view without kernels: call set_psf or set_kernels for every view");
        for (int d = 0; d < 3; ++d) {
            Reach a = reach_of(vw.k1d[d]), b = reach_of(vw.k2d[d]);
            r1[d].lo = std::max(r1[d].lo, a.lo); r1[d].hi = std::max(r1[d].hi, a.hi);
            r2[d].lo = std::max(r2[d].lo, b.lo); r2[d].hi = std::max(r2[d].hi, b.hi);
        }
    }
    for (int d = 0; d < 3; ++d) { r1_[d] = r1[d]; r2_[d] = r2[d]; }
    const bool two = cfg_.exchange_scheme == 1;
    auto need = [&](int d, bool lower) {
        const int a = lower ? r1[d].lo : r1[d].hi, b = lower ? r2[d].lo : r2[d].hi;
        return two ? std::max(a, b) : a + b;
    };
    halo_lo_ = cfg_.geom.own_lo[2] == 0 ? 0 : need(2, true);
    halo_hi_ = cfg_.geom.own_hi[2] == cfg_.geom.gdim[2] ? 0 : need(2, false);
    if (cfg_.geom.own_lo[2] - halo_lo_ < cfg_.geom.goff[2] || cfg_.geom.own_hi[2] + halo_hi_ > cfg_.geom.goff[2] + cfg_.geom.vol[2])
        throw Error("sharded context: the local arrays do not contain the halo planes (mvd_halo_planes)");
    halo_y_lo_ = cfg_.geom.own_lo[1] == 0 ? 0 : need(1, true);
    halo_y_hi_ = cfg_.geom.own_hi[1] == cfg_.geom.gdim[1] ? 0 : need(1, false);
    if (cfg_.geom.own_lo[1] - halo_y_lo_ < cfg_.geom.goff[1] || cfg_.geom.own_hi[1] + halo_y_hi_ > cfg_.geom.goff[1] + cfg_.geom.vol[1])
        throw Error("sharded context: the local arrays do not contain the halo rows (mvd_halo_rows)");
    for (View& vw : views_) { dev::free_(vw.k1hat); dev::free_(vw.k2hat); vw.k1hat = vw.k2hat = nullptr; }
    conv_.reset(new Convolver(cfg_.geom, r1, r2, 0, cfg_.max_len, stream_, tables_.get(), two));
    if (two && (sharded(1) || sharded(2))) {
        // the quotient spectrum is exchanged tile by tile: tiles may follow each other along x (whole spectrum rows travel), but the
        // box must fit one tile in y and z
        for (const TileGeom& t : conv_->tiles())
            if (t.org[1] != conv_->tiles()[0].org[1] || t.org[2] != conv_->tiles()[0].org[2])
                throw Error("exchange scheme 1 needs the local box to fit one FFT tile in y and z (use scheme 0 or more ranks)");
    }
    install_mid_exchange();
    for (View& vw : views_) {
        vw.k1hat = conv_->build_khat(vw.k1.data(), vw.k1d);
        vw.k2hat = conv_->build_khat(vw.k2.data(), vw.k2d);
    }
    const size_t nparts = (size_t)conv_->num_tiles() * conv_->parts_per_tile();
    dev::free_(part_sum_); dev::free_(part_max_);
    part_sum_ = (double*)dev::alloc(sizeof(double) * (nparts + 256));
    part_max_ = (float*)dev::alloc(sizeof(float) * (nparts + 256));
    pfor((long long)nparts + 256, FillParts{part_sum_, part_max_}, stream_);
    inited_ = true;
}

void Engine::get_kernel_dims(int v, int which, int kd[3]) const {
    const View& vw = views_.at(v);
    for (int d = 0; d < 3; ++d) kd[d] = which == 1 ? vw.k1d[d] : vw.k2d[d];
}
void Engine::get_kernel(int v, int which, float* out) const {
    const View& vw = views_.at(v);
    const std::vector<float>& k = which == 1 ? vw.k1 : vw.k2;
    std::copy(k.begin(), k.end(), out);
}

void Engine::set_psi_host(const float* psi) {
    join_halo();
    dev::set_device(cfg_.device);
    dev::h2d(psi_[cur_], psi, sizeof(float) * local_voxels(), stream_);
    dev::sync(stream_);
}
void Engine::get_psi_host(float* psi) {
    join_halo();
    dev::set_device(cfg_.device);
    dev::d2h(psi, psi_[cur_], sizeof(float) * local_voxels(), stream_);
    dev::sync(stream_);
}

int Engine::launches_per_view_update() const { return conv_ ? conv_->launches_per_update() + 2 : 0; }

void Engine::ensure_stats_slot() {
    if (stats_count_ >= stats_cap_) {
        // grow the statistics ring (keeps earlier entries)
        const int ncap = stats_cap_ ? stats_cap_ * 2 : 1024;
        double* n = (double*)dev::alloc(sizeof(double) * 2 * ncap);
        if (stats_dev_) { dev::d2d(n, stats_dev_, sizeof(double) * 2 * stats_count_, stream_); dev::sync(stream_); dev::free_(stats_dev_); }
        stats_dev_ = n;
        stats_cap_ = ncap;
    }
}

// ------------------------------------------------------------------------------------------------
// PsiInit on the device (M/process/deconvolution/init/PsiInitBlurredFused.java:63-127, PsiInitAvgPrecise.java:52-112,
// PsiInitAvgApprox.java:47-99)
// ------------------------------------------------------------------------------------------------
void Engine::all_reduce(double* values, int count, int op) {
    if (!is_sharded() || count <= 0) return;
    if (host_reduce_) {
        if (host_reduce_(host_reduce_user_, values, count, op) != 0) throw Error("the host's reduce callback failed");
    } else if (comm_) {
        comm_->all_reduce(values, count, op);
    } else {
        throw Error("a sharded context needs an attached communicator (mvd_comm_attach) or a reduce callback (mvd_set_reduce_callback) for its "
                    "global statistics");
    }
}

// Gauss3.gauss(sigma, extendMirrorSingle(psi), psi) on a sharded context: the blur reads r = k/2 voxels beyond the own box, more than the
// halo the iterations need, so the un-blurred fused estimate of the own box is copied into a temporary array with an r-wide halo, the halo
// is filled from the neighbours (their own boxes hold their part of the same global estimate), and the blurred own box is copied back.
void Engine::psi_blur_sharded(const std::vector<float>& k3, int k) {
    const Geometry& g = cfg_.geom;
    const int r = k / 2;
    Geometry gt = g;
    for (int d = 1; d < 3; ++d) {
        if (!sharded(d)) continue;
        if (g.own_hi[d] - g.own_lo[d] < r) throw Error("sharded FUSED_BLURRED: every box must be at least (int)(3*sigma+0.5) voxels thick");
        const int lo = std::max(0, g.own_lo[d] - r), hi = std::min(g.gdim[d], g.own_hi[d] + r);
        gt.goff[d] = lo; gt.vol[d] = hi - lo;
    }
    const size_t nt = (size_t)gt.vol[0] * gt.vol[1] * gt.vol[2];
    float* tin = (float*)dev::alloc(sizeof(float) * nt);
    float* tout = nullptr;
    cpx* khat = nullptr;
    try {
        tout = (float*)dev::alloc(sizeof(float) * nt);
        dev::zero(tin, sizeof(float) * nt, stream_);
        const int rows = g.own_hi[1] - g.own_lo[1], planes = g.own_hi[2] - g.own_lo[2];
        copy_region(stream_, psi_[cur_], g.vol[1], g.own_lo[1] - g.goff[1], g.own_lo[2] - g.goff[2],
                    tin, gt.vol[1], g.own_lo[1] - gt.goff[1], g.own_lo[2] - gt.goff[2], g.vol[0], rows, planes);
        HaloBox b;
        b.base = tin; b.row_floats = gt.vol[0]; b.nrows = gt.vol[1]; b.nplanes = gt.vol[2];
        b.y0 = g.own_lo[1] - gt.goff[1]; b.y1 = g.own_hi[1] - gt.goff[1];
        b.z0 = g.own_lo[2] - gt.goff[2]; b.z1 = g.own_hi[2] - gt.goff[2];
        b.hy_lo = b.hy_hi = sharded(1) ? r : 0;
        b.hz_lo = b.hz_hi = sharded(2) ? r : 0;
        do_exchange(2, b, true);
        Reach r1[3], r2[3];
        const int kd[3] = {k, k, k};
        for (int d = 0; d < 3; ++d) { r1[d] = reach_of(k); r2[d] = Reach{0, 0}; }
        Convolver cv(gt, r1, r2, 0, cfg_.max_len, stream_, tables_.get());
        khat = cv.build_khat(k3.data(), kd);
        cv.conv(tin, tout, khat, EXT_MIRROR, 0.f);
        copy_region(stream_, tout, gt.vol[1], g.own_lo[1] - gt.goff[1], g.own_lo[2] - gt.goff[2],
                    psi_[cur_], g.vol[1], g.own_lo[1] - g.goff[1], g.own_lo[2] - g.goff[2], g.vol[0], rows, planes);
        dev::sync(stream_);
    } catch (...) {
        dev::free_(tin); dev::free_(tout); dev::free_(khat);
        throw;
    }
    dev::free_(tin); dev::free_(tout); dev::free_(khat);
    exchange_psi(psi_[cur_]);                 // the halo of the iterations' psi array
}

void Engine::psi_init(int type, double sigma, double* avg_out, float* max_out, bool set_img_to_avg) {
    join_halo();
    dev::set_device(cfg_.device);
    const int V = cfg_.num_views;
    if (V > MVD_MAX_VIEWS) throw Error("too many views for the device PsiInit");
    const Geometry& g = cfg_.geom;
    const bool shard = is_sharded();
    if (shard && !can_reduce())
        throw Error("PsiInit on a sharded context needs global statistics: attach a communicator (mvd_comm_attach) or a reduce callback "
                    "(mvd_set_reduce_callback) first");
    if (shard && type == PSI_FUSED_BLURRED && !has_exchange())
        throw Error("PsiInit FUSED_BLURRED on a sharded context needs an attached exchange (mvd_comm_attach / mvd_set_exchange_callback)");
    const long long n = (long long)local_voxels();
    const OwnBox ob{g.vol[0], g.vol[1], g.own_lo[1] - g.goff[1], g.own_hi[1] - g.goff[1], g.own_lo[2] - g.goff[2], g.own_hi[2] - g.goff[2]};
    ViewPtrs vp;
    for (int j = 0; j < V; ++j) {
        if (!views_[j].img) throw Error("view without image");
        wait_upload(views_[j]);
        vp.img[j] = views_[j].img;
        vp.weight[j] = views_[j].weight;
    }
    if (!acc_dev_) acc_dev_ = (double*)dev::alloc(sizeof(double) * 2);
    if (!max_dev_) max_dev_ = (float*)dev::alloc(sizeof(float) * MVD_MAX_VIEWS);
    double acc[2] = {0, 0};
    float mx[MVD_MAX_VIEWS] = {0};
    double avg = 0, avg_reported = 0;
    auto reduce_max = [&](float lowest) {      // per-view maxima over all boxes (MultiViewDeconvolution.java:115-135: max[] is global)
        double m[MVD_MAX_VIEWS];
        for (int j = 0; j < V; ++j) m[j] = (double)mx[j];
        (void)lowest;
        all_reduce(m, V, 1);
        for (int j = 0; j < V; ++j) mx[j] = (float)m[j];
    };
    if (type == PSI_FUSED_BLURRED || type == PSI_AVG) {
        if (type == PSI_FUSED_BLURRED)
            for (int j = 0; j < V; ++j) if (!views_[j].weight) throw Error("FUSED_BLURRED needs the view weights");
        psi_fused_stats(stream_, vp, V, type == PSI_FUSED_BLURRED ? psi_[cur_] : nullptr, n, ob, acc_dev_, max_dev_);
        dev::d2h(acc, acc_dev_, sizeof(acc), stream_);
        dev::d2h(mx, max_dev_, sizeof(float) * V, stream_);
        dev::sync(stream_);
        all_reduce(acc, 2, 0);
        reduce_max(0.f);
        if (acc[1] == 0 && type == PSI_FUSED_BLURRED)        // PsiInitBlurredFused.java:95-99; PsiInitAvgPrecise has no such failure (:90-97: NaN -> 1)
            throw Error("None of the views covers the deconvolved area, did you set the bounding box right?");
        avg = acc[0] / acc[1];
        if (avg != avg) avg = 1.0;
        avg_reported = avg;
        if (type == PSI_AVG) {
            if (set_img_to_avg) fill_volume(stream_, psi_[cur_], n, (float)avg);
        } else {
            // Gauss3.gauss(sigma, extendMirrorSingle(psi), psi) -- separable Gaussian expressed as one 3-d kernel through the FFT passes
            const std::vector<double> half = gauss3_halfkernel(sigma);
            const int r = (int)half.size() - 1, k = 2 * r + 1;
            std::vector<float> k3((size_t)k * k * k);
            for (int z = 0; z < k; ++z)
                for (int y = 0; y < k; ++y)
                    for (int x = 0; x < k; ++x)
                        k3[((size_t)z * k + y) * k + x] = (float)(half[std::abs(z - r)] * half[std::abs(y - r)] * half[std::abs(x - r)]);
            if (shard) {
                psi_blur_sharded(k3, k);
            } else {
                Reach r1[3], r2[3];
                const int kd[3] = {k, k, k};
                for (int d = 0; d < 3; ++d) { r1[d] = reach_of(k); r2[d] = Reach{0, 0}; }
                Convolver cv(g, r1, r2, 0, cfg_.max_len, stream_, tables_.get());
                cpx* khat = cv.build_khat(k3.data(), kd);
                cv.conv(psi_[cur_], psi_[cur_ ^ 1], khat, EXT_MIRROR, 0.f);
                dev::sync(stream_);
                dev::free_(khat);
                cur_ ^= 1;
            }
        }
    } else if (type == PSI_APPROX_AVG) {
        // PsiInitAvgApproxThread.java:58-85: min / max / mean of the central x-hyperslice (visited numDimensions times: same mean)
        double sums[2 * MVD_MAX_VIEWS];
        for (int j = 0; j < V; ++j) {
            dev::zero(acc_dev_, sizeof(double) * 2, stream_);
            const float lowest = -3.0e38f;
            dev::h2d(max_dev_, &lowest, sizeof(float), stream_);
            dev::sync(stream_);
            slice_stats(stream_, views_[j].img, ob, (long long)g.vol[1] * g.vol[2], acc_dev_, max_dev_);
            dev::d2h(acc, acc_dev_, sizeof(acc), stream_);
            dev::d2h(&mx[j], max_dev_, sizeof(float), stream_);
            dev::sync(stream_);
            sums[2 * j] = acc[0]; sums[2 * j + 1] = acc[1];
        }
        all_reduce(sums, 2 * V, 0);
        reduce_max(-3.0e38f);
        for (int j = 0; j < V; ++j) avg += sums[2 * j] / sums[2 * j + 1];
        avg /= (double)V;
        if (avg != avg) avg = 1.0;
        avg_reported = -1.0;                                 // PsiInitAvgApprox.getAvg(): the field is shadowed by a local (:40,57,80)
        if (set_img_to_avg) fill_volume(stream_, psi_[cur_], n, (float)avg);
    } else {
        throw Error("unknown PsiInit type");
    }
    for (int j = 0; j < V; ++j) views_[j].max_intensity = mx[j];
    dev::sync(stream_);
    if (avg_out) *avg_out = avg_reported;
    if (max_out) for (int j = 0; j < V; ++j) max_out[j] = mx[j];
}

void Engine::make_blending_weights(int v, const int box_min[3], const int box_max[3], const float border[3], const float blending[3],
                                   const double* inv_affine, const int* bbox_offset) {
    if (v < 0 || v >= cfg_.num_views) throw Error("view index out of range");
    dev::set_device(cfg_.device);
    View& vw = views_[v];
    wait_upload(vw);          // an asynchronous upload of this view zeroes / fills weight_owned on the copy stream
    if (vw.weight && !vw.weight_owned) throw Error("the weight of this view is borrowed device memory; cannot generate into it");
    if (!vw.weight_owned) vw.weight_owned = (float*)dev::alloc(sizeof(float) * local_voxels());
    vw.weight = vw.weight_owned;
    weights_changed();
    if (!lut_dev_) {
        const std::vector<double> lut = blend_lut();
        lut_dev_ = (double*)dev::alloc(sizeof(double) * lut.size());
        dev::h2d(lut_dev_, lut.data(), sizeof(double) * lut.size(), stream_);
        dev::sync(stream_);
    }
    blend_weights(stream_, vw.weight_owned, lut_dev_, cfg_.geom.vol, cfg_.geom.goff, box_min, box_max, border, blending, inv_affine, bbox_offset);
}

void Engine::fuse_group_host(int v, const RawViewDev* views_host, int count, const int bbox_min[3], float min_value_img, float outside_value) {
    if (v < 0 || v >= cfg_.num_views) throw Error("view index out of range");
    if (count < 1 || count > 64) throw Error("a group needs 1..64 views");
    dev::set_device(cfg_.device);
    View& vw = views_[v];
    wait_upload(vw);
    if ((vw.img && !vw.img_owned) || (vw.weight && !vw.weight_owned)) throw Error("this view borrows device memory; cannot generate into it");
    const size_t bytes = sizeof(float) * local_voxels();
    if (!vw.img_owned) vw.img_owned = (float*)dev::alloc(bytes);
    if (!vw.weight_owned) vw.weight_owned = (float*)dev::alloc(bytes);
    if (!lut_dev_) {
        const std::vector<double> lut = blend_lut();
        lut_dev_ = (double*)dev::alloc(sizeof(double) * lut.size());
        dev::h2d(lut_dev_, lut.data(), sizeof(double) * lut.size(), stream_);
    }
    std::vector<RawViewDev> hv(views_host, views_host + count);
    std::vector<float*> raws;
    RawViewDev* dv = nullptr;
    try {
        for (RawViewDev& r : hv) {
            for (int d = 0; d < 3; ++d) if (r.dims[d] < 2) throw Error("raw views need at least 2 samples per axis");
            if (!r.raw) throw Error("null raw view");
            const size_t n = (size_t)r.dims[0] * r.dims[1] * r.dims[2];
            float* p = (float*)dev::alloc(sizeof(float) * n);
            raws.push_back(p);
            dev::h2d(p, r.raw, sizeof(float) * n, stream_);
            r.raw = p;
        }
        dv = (RawViewDev*)dev::alloc(sizeof(RawViewDev) * hv.size());
        dev::h2d(dv, hv.data(), sizeof(RawViewDev) * hv.size(), stream_);
#ifndef MVD_HOST_EMU
        cudaEvent_t e0, e1;
        MVD_CUDA_CHECK(cudaEventCreate(&e0)); MVD_CUDA_CHECK(cudaEventCreate(&e1));
        MVD_CUDA_CHECK(cudaEventRecord(e0, stream_));
#endif
        fuse_group(stream_, dv, hv.data(), count, vw.img_owned, vw.weight_owned, lut_dev_, cfg_.geom.vol, cfg_.geom.goff, bbox_min, min_value_img, outside_value);
#ifndef MVD_HOST_EMU
        MVD_CUDA_CHECK(cudaEventRecord(e1, stream_));
#endif
        dev::sync(stream_);
#ifndef MVD_HOST_EMU
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        last_fuse_ms_ = ms;
#endif
    } catch (...) {
        for (float* p : raws) dev::free_(p);
        dev::free_(dv);
        throw;
    }
    for (float* p : raws) dev::free_(p);
    dev::free_(dv);
    vw.img = vw.img_owned;
    vw.weight = vw.weight_owned;
    weights_changed();
}

void Engine::normalize_view_weights(double osem_speedup, bool additional_smooth, float max_diff_range, float scaling_range) {
    dev::set_device(cfg_.device);
    const int V = cfg_.num_views;
    if (V > MVD_MAX_VIEWS) throw Error("too many views");
    WeightPtrs w;
    for (int j = 0; j < V; ++j) {
        if (!views_[j].weight_owned || views_[j].weight != views_[j].weight_owned) throw Error("normalisation needs context-owned weights for every view");
        wait_upload(views_[j]);
        w.w[j] = views_[j].weight_owned;
    }
    normalize_weights(stream_, w, V, (long long)local_voxels(), osem_speedup, additional_smooth, max_diff_range, scaling_range);
    weights_changed();
}

void Engine::get_image_host(int v, float* out) {
    if (v < 0 || v >= cfg_.num_views || !views_[v].img) throw Error("no such image");
    dev::set_device(cfg_.device);
    View& vw = views_[v];
    wait_upload(vw);
    dev::d2h(out, vw.img, sizeof(float) * local_voxels(), stream_);
    dev::sync(stream_);
}
void Engine::get_weight_host(int v, float* out) {
    if (v < 0 || v >= cfg_.num_views || !views_[v].weight) throw Error("no such weight");
    dev::set_device(cfg_.device);
    wait_upload(views_[v]);
    dev::d2h(out, views_[v].weight, sizeof(float) * local_voxels(), stream_);
    dev::sync(stream_);
}

// MultiViewDeconvolutionMul.runNextIteration / ComputeBlockMulThreadCPU.runIteration (mul/ComputeBlockMulThreadCPU.java:87-188)
void Engine::iteration_mul() {
    join_halo();
    if (!inited_) throw Error("init_views() has not been called");
    dev::set_device(cfg_.device);
    const int V = cfg_.num_views;
    if (V > MVD_MAX_VIEWS) throw Error("too many views");
    const Geometry& g = cfg_.geom;
    if (g.own_lo[1] != 0 || g.own_hi[1] != g.gdim[1]) throw Error("the Mul iteration supports z-slab sharding only");
    if (cfg_.exchange_scheme == 1 && sharded(2)) throw Error("the Mul iteration needs exchange scheme 0");
    const long long plane = (long long)g.vol[0] * g.vol[1], n = (long long)local_voxels();
    const long long own0 = (long long)(g.own_lo[2] - g.goff[2]) * plane, own1 = (long long)(g.own_hi[2] - g.goff[2]) * plane;
    if ((int)integral_.size() < V) {
        integral_.resize(V, nullptr);
        for (float*& p : integral_) if (!p) p = (float*)dev::alloc(sizeof(float) * n);
    }
    if (!max_dev_) max_dev_ = (float*)dev::alloc(sizeof(float) * MVD_MAX_VIEWS);
    ensure_stats_slot();
    MulPtrs mp;
    double miv = 0;
    for (int v = 0; v < V; ++v) {                            // all views from the SAME psi
        View& vw = views_[v];
        if (!vw.img || !vw.weight) throw Error("view without image/weight");
        wait_upload(vw);
        conv_->integral(psi_[cur_], vw.img, vw.k1hat, vw.k2hat, integral_[v]);
        mp.integral[v] = integral_[v];
        mp.weight[v] = vw.weight;
        miv += (double)vw.max_intensity;
    }
    miv /= (double)V;                                        // :144-149
    mul_combine(stream_, mp, V, psi_[cur_], psi_[cur_ ^ 1], n, own0, own1, cfg_.lambda, cfg_.min_value, (float)miv,
                stats_dev_ + 2 * (size_t)stats_count_, max_dev_);
    ++stats_count_;
    cur_ ^= 1;
    if (has_exchange()) exchange_psi(psi_[cur_]);
}

void Engine::view_update(int v) {
    if (!inited_) throw Error("init_views() has not been called");
    if (v < 0 || v >= cfg_.num_views) throw Error("view index out of range");
    dev::set_device(cfg_.device);
    View& vw = views_[v];
    if (!vw.img || !vw.weight) throw Error("view without image/weight");
    if (cfg_.exchange_scheme == 1 && (sharded(1) || sharded(2)) && !has_exchange())
        throw Error("exchange scheme 1: attach a communicator (mvd_comm_attach) or an exchange callback before the first view update");
    wait_upload(vw);
    ensure_stats_slot();
    const int nparts = conv_->num_tiles() * conv_->parts_per_tile();
    const unsigned char* skip = nullptr;
    if (skip_on_) {
        if (!skip_valid_) refresh_skip();
        skip = skip_.data() + (size_t)v * conv_->num_tiles();
    }
    // a psi exchange that is still travelling is joined inside the update, after the lines that do not depend on it
    const std::function<void()> join = [this] { join_halo(); };
    conv_->view_update(psi_[cur_], psi_[cur_ ^ 1], vw.img, vw.weight, vw.k1hat, vw.k2hat, cfg_.lambda, cfg_.min_value,
                       vw.max_intensity, part_sum_, part_max_, skip, halo_pending_ ? &join : nullptr);
    join_halo();                                          // (every tile skipped: nothing joined it)
    cur_ ^= 1;
    if (has_exchange()) start_psi_exchange(psi_[cur_]);   // travels behind the statistics kernels and the next update's first lines
    // deterministic two-level reduction of the per-CTA partial statistics
    reduce_parts(stream_, part_sum_, part_max_, nparts, part_sum_ + nparts, part_max_ + nparts, stats_dev_ + 2 * (size_t)stats_count_);
    ++stats_count_;
}

int Engine::skip_empty_tiles(bool on) {
    if (!inited_) throw Error("init_views() has not been called");
    skip_on_ = on;
    skip_valid_ = false;
    return on ? refresh_skip() : 0;
}

int Engine::refresh_skip() {
    dev::set_device(cfg_.device);
    const Geometry& g = cfg_.geom;
    const int V = cfg_.num_views, nt = conv_->num_tiles();
    if (!flag_dev_) flag_dev_ = (int*)dev::alloc(sizeof(int) * (size_t)V * nt);
    dev::zero(flag_dev_, sizeof(int) * (size_t)V * nt, stream_);
    for (int v = 0; v < V; ++v) {
        View& vw = views_[v];
        if (!vw.weight) throw Error("view without image/weight");
        wait_upload(vw);
        for (int ti = 0; ti < nt; ++ti) {
            const TileGeom& t = conv_->tiles()[ti];
            int lo[3], hi[3];
            for (int d = 0; d < 3; ++d) { lo[d] = t.lo[d] - g.goff[d]; hi[d] = t.hi[d] - g.goff[d]; }
            box_nonzero(stream_, vw.weight, g.vol[0], g.vol[1], lo, hi, flag_dev_ + (size_t)v * nt + ti);
        }
    }
    std::vector<int> flags((size_t)V * nt);
    dev::d2h(flags.data(), flag_dev_, sizeof(int) * flags.size(), stream_);
    dev::sync(stream_);
    if (cfg_.exchange_scheme == 1 && is_sharded()) {
        // the neighbours read this box's quotient between the two convolutions: a tile can only be dropped when no rank has content
        if (!can_reduce()) throw Error("skip_empty_tiles on a sharded scheme-1 context needs the communicator / reduce callback first");
        std::vector<double> any(flags.begin(), flags.end());
        all_reduce(any.data(), (int)any.size(), 1);
        for (size_t i = 0; i < flags.size(); ++i) flags[i] = any[i] > 0.0 ? 1 : 0;
    }
    skip_.assign(flags.size(), 0);
    int skipped = 0;
    for (size_t i = 0; i < flags.size(); ++i) { skip_[i] = flags[i] ? 0 : 1; skipped += skip_[i]; }
    skip_valid_ = true;
    return skipped;
}

void Engine::comm_attach(std::shared_ptr<NcclComm> comm, int py, int pz) {
    if (!inited_) throw Error("init_views() must run before the communicator is attached (the halo widths come from the kernels)");
    if (!comm || comm->device() != cfg_.device) throw Error("communicator belongs to another device");
    dev::set_device(cfg_.device);
    if ((py > 1) != sharded(1) || (pz > 1) != sharded(2)) throw Error("process grid does not match the sharding of this context");
    size_t need_y = 0, need_z = 0;
    auto account = [&](const HaloBox& b) {
        need_y = std::max(need_y, (size_t)std::max(b.hy_lo, b.hy_hi) * (size_t)b.row_floats * (size_t)(b.z1 - b.z0));
        need_z = std::max(need_z, (size_t)std::max(b.hz_lo, b.hz_hi) * (size_t)b.row_floats * (size_t)b.nrows);
    };
    account(psi_box(psi_[0]));
    if (cfg_.exchange_scheme == 1) account(spectrum_box(nullptr, conv_->tiles().at(0)));
    join_halo();
    comm_.reset(new HaloComm(std::move(comm), py, pz, stream_, need_y, need_z));
    // peer transport: the exchanges get their own high-priority stream and overlap the passes (MVD_OVERLAP=0: serialised on the compute stream)
    const char* ov = std::getenv("MVD_OVERLAP");
    overlap_ = comm_->transport() == 1 && !(ov && std::atoi(ov) == 0);
    if (overlap_ && !xstream_) {
        xstream_ = dev::stream_create_high_priority();
        ev_compute_ = dev::event_create(); ev_psi_ = dev::event_create(); ev_mid_ = dev::event_create();
    }
    install_mid_exchange();
}

void Engine::do_exchange(int which, const HaloBox& b, bool oversize, dev::event_t ev) {
    if (host_exchange_) {
        dev::sync(stream_);
        if (host_exchange_(host_exchange_user_, which, &b) != 0) throw Error("the host's exchange callback failed");
    } else if (comm_) {
        if (!overlap_) { comm_->exchange(b, oversize); return; }
        // every exchange of an overlapping context runs on the exchange stream (they share landing buffers and sequence numbers, so
        // they must stay ordered among themselves): it starts behind what the compute stream has enqueued so far ...
        dev::event_record(ev_compute_, stream_);
        dev::stream_wait(xstream_, ev_compute_);
        comm_->exchange(b, oversize, xstream_);
        // ... and the compute stream either continues and joins later (ev) or waits right away
        dev::event_record(ev ? ev : ev_mid_, xstream_);
        if (!ev) dev::stream_wait(stream_, ev_mid_);
    }
}

void Engine::start_psi_exchange(float* psi) {
    if (host_exchange_) { pending_psi_ = psi; halo_pending_ = true; return; }     // deferred to the join
    if (!comm_) return;
    if (!overlap_) { do_exchange(0, psi_box(psi)); return; }
    do_exchange(0, psi_box(psi), false, ev_psi_);
    halo_pending_ = true;
}

void Engine::join_halo() {
    if (!halo_pending_) return;
    halo_pending_ = false;
    if (host_exchange_) do_exchange(0, psi_box(pending_psi_));
    else dev::stream_wait(stream_, ev_psi_);
}

// psi: scheme 0 ships what both convolutions read beyond the own box (r1 + r2), scheme 1 only the first convolution's reach
HaloBox Engine::psi_box(float* psi) const {
    const Geometry& g = cfg_.geom;
    const bool two = cfg_.exchange_scheme == 1;
    HaloBox b;
    b.base = psi; b.row_floats = g.vol[0]; b.nrows = g.vol[1]; b.nplanes = g.vol[2];
    b.y0 = g.own_lo[1] - g.goff[1]; b.y1 = g.own_hi[1] - g.goff[1];
    b.z0 = g.own_lo[2] - g.goff[2]; b.z1 = g.own_hi[2] - g.goff[2];
    b.hy_lo = sharded(1) ? r1_[1].lo + (two ? 0 : r2_[1].lo) : 0; b.hy_hi = sharded(1) ? r1_[1].hi + (two ? 0 : r2_[1].hi) : 0;
    b.hz_lo = sharded(2) ? r1_[2].lo + (two ? 0 : r2_[2].lo) : 0; b.hz_hi = sharded(2) ? r1_[2].hi + (two ? 0 : r2_[2].hi) : 0;
    return b;
}
void Engine::exchange_psi(float* psi) { do_exchange(0, psi_box(psi)); }

// scheme 1: between the two convolutions the quotient of the neighbours' own boxes replaces this box's halo rows / planes.  The
// quotient only exists as its x-spectrum (P5 fuses inverse-x -> quotient -> forward-x); the x transform is per row, so the rows and
// planes of the spectrum are exchanged instead -- same geometry, rows of 2 * pitch floats.
HaloBox Engine::spectrum_box(cpx* work, const TileGeom& t) const {
    const Geometry& g = cfg_.geom;
    const int* T = conv_->tile_dims();
    HaloBox b;
    b.base = reinterpret_cast<float*>(work); b.row_floats = 2LL * conv_->pitch(); b.nrows = T[1]; b.nplanes = T[2];
    b.y0 = g.own_lo[1] - t.org[1]; b.y1 = g.own_hi[1] - t.org[1];
    b.z0 = g.own_lo[2] - t.org[2]; b.z1 = g.own_hi[2] - t.org[2];
    b.hy_lo = sharded(1) ? r2_[1].lo : 0; b.hy_hi = sharded(1) ? r2_[1].hi : 0;
    b.hz_lo = sharded(2) ? r2_[2].lo : 0; b.hz_hi = sharded(2) ? r2_[2].hi : 0;
    return b;
}
void Engine::install_mid_exchange() {
    if (!conv_) return;
    if (cfg_.exchange_scheme != 1 || !(sharded(1) || sharded(2)) || !has_exchange()) { conv_->set_mid_exchange(nullptr); return; }
    // start + join: the quotient pass runs boundary first and the exchange travels behind its deep interior (host callback: the exchange
    // is complete when start returns, the pass is split all the same)
    conv_->set_mid_exchange([this](cpx* work, const TileGeom& t) { do_exchange(1, spectrum_box(work, t), false, overlap_ ? ev_mid_ : nullptr); },
                            [this] { if (overlap_) dev::stream_wait(stream_, ev_mid_); });
}

void Engine::exchange_halos() {
    join_halo();
    if (!has_exchange()) throw Error("no communicator attached (mvd_comm_attach / mvd_set_exchange_callback)");
    dev::set_device(cfg_.device);
    exchange_psi(psi_[cur_]);
}

void Engine::fetch_stats(int count, IterStats* out) {
    join_halo();
    dev::set_device(cfg_.device);
    if (count > stats_count_) count = stats_count_;
    std::vector<double> h(2 * (size_t)std::max(count, 1));
    if (count > 0) dev::d2h(h.data(), stats_dev_ + 2 * (size_t)(stats_count_ - count), sizeof(double) * 2 * count, stream_);
    dev::sync(stream_);
    if (is_sharded() && can_reduce() && count > 0) {
        // IterationStatistics over the whole volume (MultiViewDeconvolutionSeq.java:165-176 sums the blocks' sumChange and takes the largest
        // maxChange): collective -- every rank fetches the same number of entries
        std::vector<double> sums((size_t)count), maxs((size_t)count);
        for (int i = 0; i < count; ++i) { sums[i] = h[2 * i]; maxs[i] = h[2 * i + 1]; }
        all_reduce(sums.data(), count, 0);
        all_reduce(maxs.data(), count, 1);
        for (int i = 0; i < count; ++i) { h[2 * i] = sums[i]; h[2 * i + 1] = maxs[i]; }
    }
    for (int i = 0; i < count; ++i) { out[i].sum_change = h[2 * i]; out[i].max_change = h[2 * i + 1]; }
}

void Engine::run_iterations(int n, IterStats* out) {
    stats_count_ = 0;
    for (int it = 0; it < n; ++it)
        for (int v = 0; v < cfg_.num_views; ++v) view_update(v);      // OSEM: psi updated after every view
    if (out) fetch_stats(n * cfg_.num_views, out);
    else { join_halo(); dev::sync(stream_); }
}

void convolve_host(int device, stream_t stream_, Tables* tables, int max_len, const float* src, const int dims[3],
                   const float* kernel, const int kd[3], int ext, float ext_value, float* dst, bool circular) {
    dev::set_device(device);
    Geometry g;
    Reach r1[3], r2[3];
    for (int d = 0; d < 3; ++d) {
        g.gdim[d] = g.vol[d] = dims[d];
        g.goff[d] = 0; g.own_lo[d] = 0; g.own_hi[d] = dims[d];
        r1[d] = reach_of(kd[d]);
        r2[d] = Reach{0, 0};
    }
    Convolver cv(g, r1, r2, circular ? 1 : 0, max_len, stream_, tables);
    const size_t n = (size_t)dims[0] * dims[1] * dims[2];
    float* s = (float*)dev::alloc(sizeof(float) * n);
    float* d_ = (float*)dev::alloc(sizeof(float) * n);
    cpx* khat = nullptr;
    try {
        dev::h2d(s, src, sizeof(float) * n, stream_);
        khat = cv.build_khat(kernel, kd);
        cv.conv(s, d_, khat, ext, ext_value);
        dev::d2h(dst, d_, sizeof(float) * n, stream_);
        dev::sync(stream_);
    } catch (...) {
        dev::free_(s); dev::free_(d_); dev::free_(khat);
        throw;
    }
    dev::free_(s); dev::free_(d_); dev::free_(khat);
}

// ------------------------------------------------------------------------------------------------
// per-pass event timing: an event is recorded in front of every pass launch (and one behind the last); the time between
// consecutive events is attributed to the pass launched in between.
// ------------------------------------------------------------------------------------------------
void Convolver::mark(int pass) {
#ifndef MVD_HOST_EMU
    if (!prof_on_) return;
    if (prof_used_ == prof_events_.size()) {
        cudaEvent_t e;
        MVD_CUDA_CHECK(cudaEventCreate(&e));
        prof_events_.push_back((void*)e);
        prof_ids_.push_back(-1);
    }
    MVD_CUDA_CHECK(cudaEventRecord((cudaEvent_t)prof_events_[prof_used_], stream_));
    prof_ids_[prof_used_] = pass;
    ++prof_used_;
#else
    (void)pass;
#endif
}
void Convolver::collect_pass_times(double ms[9], long long counts[9], bool reset) {
#ifndef MVD_HOST_EMU
    dev::sync(stream_);
    for (size_t i = 0; i + 1 < prof_used_; ++i) {
        const int id = prof_ids_[i];
        if (id < 0 || id > 11) continue;
        float t = 0.f;
        MVD_CUDA_CHECK(cudaEventElapsedTime(&t, (cudaEvent_t)prof_events_[i], (cudaEvent_t)prof_events_[i + 1]));
        prof_ms_[id] += t;
        prof_n_[id] += 1;
    }
    prof_used_ = 0;
#endif
    for (int i = 0; i < 9; ++i) { ms[i] = prof_ms_[i]; counts[i] = prof_n_[i]; }
    if (reset) for (int i = 0; i < 9; ++i) { prof_ms_[i] = 0; prof_n_[i] = 0; }
}
void Convolver::collect_aux_times(double ms[3], long long counts[3], bool reset) {
    double pm[9]; long long pn[9];
    collect_pass_times(pm, pn, false);
    for (int i = 0; i < 3; ++i) { ms[i] = prof_ms_[9 + i]; counts[i] = prof_n_[9 + i]; }
    if (reset) for (int i = 9; i < 12; ++i) { prof_ms_[i] = 0; prof_n_[i] = 0; }
}

}  // namespace mvd
