// Point-wise device stages around the FFT passes: PsiInit (fused / average initial estimate + per-view maxima), weight generation
// (cosine blending of a view's box) and normalisation, and the combination step of the non-OSEM ("Mul") iteration.
// Compiled by nvcc for the product and by g++ (-DMVD_HOST_EMU) for the CPU-side tests.
#include "engine.h"

#include <algorithm>
#include <cmath>

namespace mvd {

#ifndef MVD_HOST_EMU
namespace {
__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_down_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {     // signed-safe
    int* a = reinterpret_cast<int*>(addr);
    int old = *a;
    while (__int_as_float(old) < v) {
        const int assumed = old;
        old = atomicCAS(a, assumed, __float_as_int(v));
        if (old == assumed) break;
    }
}
// block-level reduction of (sum, count, max) followed by one atomic per block
__device__ void block_accumulate(double s, double c, float m, double* gsum, double* gcount, float* gmax) {
    __shared__ double ss[32], sc[32];
    __shared__ float sm_[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    s = warp_sum(s); c = warp_sum(c); m = warp_max(m);
    if (lane == 0) { ss[wid] = s; sc[wid] = c; sm_[wid] = m; }
    __syncthreads();
    if (wid == 0) {
        s = lane < nw ? ss[lane] : 0.0; c = lane < nw ? sc[lane] : 0.0; m = lane < nw ? sm_[lane] : -3.0e38f;
        s = warp_sum(s); c = warp_sum(c); m = warp_max(m);
        if (lane == 0) {
            if (gsum) atomicAdd(gsum, s);
            if (gcount) atomicAdd(gcount, c);
            if (gmax) atomic_max_float(gmax, m);
        }
    }
    __syncthreads();
}
}  // namespace
#endif

// ---------------------------------------------------------------------------------------------------------------------
// FusedNonZeroRandomAccess.get over the local array + the statistics of PsiInitBlurredFused / PsiInitAvgPrecise
// (M/process/deconvolution/util/FusedNonZeroRandomAccess.java:57-96, init/PsiInitBlurredFused.java:76-101,
//  init/PsiInitAvgPreciseThread.java:127-155).  acc = {sum of per-voxel mean positive intensity, #covered voxels}; statistics only over
// the owned index range [own0, own1).
// ---------------------------------------------------------------------------------------------------------------------
MVD_HD float fused_voxel(const ViewPtrs& vp, int V, long long i, double& mean_pos, bool& covered, float* vmax /*[V] or null*/) {
    double sumI = 0, sumW = 0, sum = 0;
    int count = 0;
    for (int j = 0; j < V; ++j) {
        const double intensity = (double)vp.img[j][i];
        if (intensity > 0) {
            const double weight = vp.weight[j] ? (double)vp.weight[j][i] : 1.0;
            sumI += intensity * weight;
            sumW += weight;
            if (vmax) vmax[j] = vmax[j] > (float)intensity ? vmax[j] : (float)intensity;
            sum += intensity;
            ++count;
        }
    }
    covered = count > 0;
    mean_pos = covered ? sum / count : 0.0;
    return sumW > 0 ? (float)(sumI / sumW) : 0.f;
}

MVD_HD bool in_own_box(const OwnBox& b, long long i) {
    const long long r = i / b.nx;
    const int y = (int)(r % b.ny), z = (int)(r / b.ny);
    return y >= b.y0 && y < b.y1 && z >= b.z0 && z < b.z1;
}

#ifndef MVD_HOST_EMU
__global__ void psi_fused_kernel(ViewPtrs vp, int V, float* psi, long long n, OwnBox ob, double* acc, float* gmax) {
    double s = 0, c = 0;
    float vmax[MVD_MAX_VIEWS];
    for (int j = 0; j < MVD_MAX_VIEWS; ++j) vmax[j] = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double mp; bool cov;
        const bool own = in_own_box(ob, i);
        const float v = fused_voxel(vp, V, i, mp, cov, own ? vmax : nullptr);
        if (psi) psi[i] = v;
        if (own && cov) { s += mp; c += 1.0; }
    }
    block_accumulate(s, c, 0.f, acc, acc + 1, nullptr);
    for (int j = 0; j < V; ++j) block_accumulate(0, 0, vmax[j], nullptr, nullptr, gmax + j);
}
#endif

void psi_fused_stats(stream_t s, const ViewPtrs& vp, int V, float* psi, long long n, const OwnBox& ob, double* acc_dev, float* max_dev) {
    dev::zero(acc_dev, sizeof(double) * 2, s);
    dev::zero(max_dev, sizeof(float) * MVD_MAX_VIEWS, s);
#ifndef MVD_HOST_EMU
    psi_fused_kernel<<<148 * 8, 256, 0, s>>>(vp, V, psi, n, ob, acc_dev, max_dev);
    MVD_CUDA_CHECK(cudaGetLastError());
#else
    double sum = 0, cnt = 0;
    for (long long i = 0; i < n; ++i) {
        double mp; bool cov;
        const bool own = in_own_box(ob, i);
        const float v = fused_voxel(vp, V, i, mp, cov, own ? max_dev : nullptr);
        if (psi) psi[i] = v;
        if (own && cov) { sum += mp; cnt += 1.0; }
    }
    acc_dev[0] = sum; acc_dev[1] = cnt;
#endif
}

struct FillValue {
    float* p; float v;
    MVD_HD void operator()(long long i) const { p[i] = v; }
};
void fill_volume(stream_t s, float* p, long long n, float v) { pfor(n, FillValue{p, v}, s); }

// PsiInitAvgApproxThread (init/PsiInitAvgApproxThread.java:58-85): min / max / mean of the central x-hyperslice
struct SliceStats {
    const float* img; OwnBox ob; long long nyz; double* acc; float* gmax;   // acc = {sum, count}
};
#ifndef MVD_HOST_EMU
__global__ void slice_stats_kernel(SliceStats a) {
    double s = 0, c = 0; float m = -3.0e38f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.nyz; i += (long long)gridDim.x * blockDim.x) {
        if (!in_own_box(a.ob, i * a.ob.nx)) continue;
        const float v = a.img[i * a.ob.nx + a.ob.nx / 2];
        s += (double)v; c += 1.0; m = fmaxf(m, v);
    }
    block_accumulate(s, c, m, a.acc, a.acc + 1, a.gmax);
}
#endif
void slice_stats(stream_t s, const float* img, const OwnBox& ob, long long nyz, double* acc_dev, float* max_dev) {
    SliceStats a{img, ob, nyz, acc_dev, max_dev};
#ifndef MVD_HOST_EMU
    slice_stats_kernel<<<148, 256, 0, s>>>(a);
    MVD_CUDA_CHECK(cudaGetLastError());
#else
    double sum = 0, cnt = 0; float m = -3.0e38f;
    for (long long i = 0; i < nyz; ++i) {
        if (!in_own_box(ob, i * ob.nx)) continue;
        const float v = img[i * ob.nx + ob.nx / 2]; sum += v; cnt += 1.0; m = std::max(m, v);
    }
    acc_dev[0] += sum; acc_dev[1] += cnt; *max_dev = std::max(*max_dev, m);
#endif
}

// copy rows x planes of nx floats between two [.][ny][nx] arrays of different extents (PsiInit on a sharded context)
struct CopyRegion {
    const float* src; float* dst; int nx, rows; int sny, sy0, sz0, dny, dy0, dz0;
    MVD_HD void operator()(long long i) const {
        const int x = (int)(i % nx);
        const long long r = i / nx;
        const int y = (int)(r % rows), z = (int)(r / rows);
        dst[((long long)(dz0 + z) * dny + (dy0 + y)) * nx + x] = src[((long long)(sz0 + z) * sny + (sy0 + y)) * nx + x];
    }
};
void copy_region(stream_t s, const float* src, int sny, int sy0, int sz0, float* dst, int dny, int dy0, int dz0, int nx, int rows, int planes) {
    pfor((long long)nx * rows * planes, CopyRegion{src, dst, nx, rows, sny, sy0, sz0, dny, dy0, dz0}, s);
}

// DeconView.blockContainsContent (M/process/deconvolution/DeconView.java:232-274): does the weight volume hold a value != 0 inside a box
// (local coordinates, half open)?  flag is set to 1 by every thread that finds one (the same value from all writers).
struct BoxNonZero {
    const float* w; int nx, ny; int x0, y0, z0, bx, by; int* flag;
    MVD_HD void operator()(long long i) const {
        const int x = (int)(i % bx);
        const long long r = i / bx;
        const int y = (int)(r % by), z = (int)(r / by);
        if (w[((long long)(z0 + z) * ny + (y0 + y)) * nx + (x0 + x)] != 0.f) *flag = 1;
    }
};
void box_nonzero(stream_t s, const float* w, int nx, int ny, const int lo[3], const int hi[3], int* flag_dev) {
    const int bx = hi[0] - lo[0], by = hi[1] - lo[1], bz = hi[2] - lo[2];
    if (bx <= 0 || by <= 0 || bz <= 0) return;
    pfor((long long)bx * by * bz, BoxNonZero{w, nx, ny, lo[0], lo[1], lo[2], bx, by, flag_dev}, s);
}
struct NeutralParts {
    double* ps; float* pm;
    MVD_HD void operator()(long long i) const { ps[i] = 0.0; pm[i] = -1.f; }
};
void clear_parts(stream_t s, double* part_sum, float* part_max, int n) { pfor(n, NeutralParts{part_sum, part_max}, s); }

// ---------------------------------------------------------------------------------------------------------------------
// BlendingRealRandomAccess.computeWeight (M/process/fusion/transformed/weights/BlendingRealRandomAccess.java:95-130) for an
// axis-aligned box on the integer grid; lut = the 1001-entry cosine table built exactly like the reference's static initialiser.
// ---------------------------------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
MVD_HD double d_mul(double a, double b) { return __dmul_rn(a, b); }
MVD_HD double d_add(double a, double b) { return __dadd_rn(a, b); }
#else
MVD_HD double d_mul(double a, double b) { volatile double r = a * b; return r; }
MVD_HD double d_add(double a, double b) { volatile double r = a + b; return r; }
#endif

struct BlendKernel {
    float* out; const double* lut;
    int nx, ny; int goff[3];
    int mn[3], dim_minus1[3];
    float border[3], blending[3];
    int affine;            // 1: location = inverse affine of (voxel + offset)  (TransformedRasteredRandomAccess.applyInverse, :96-114)
    double im[12];         // row-packed inverse of the view -> fused-space transform
    int offset[3];         // bounding-box min (the fused volume is zero-min)
    MVD_HD void operator()(long long i) const {
        const int x = (int)(i % nx), y = (int)((i / nx) % ny), z = (int)(i / ((long long)nx * ny));
        float loc[3] = {(float)(x + goff[0]), (float)(y + goff[1]), (float)(z + goff[2])};
        if (affine) {
            const double t0 = (double)(x + goff[0] + offset[0]), t1 = (double)(y + goff[1] + offset[1]), t2 = (double)(z + goff[2] + offset[2]);
            for (int r = 0; r < 3; ++r)      // s = t0*i0 + t1*i1 + t2*i2 + i3, evaluated left to right in double without contraction
                loc[r] = (float)d_add(d_add(d_add(d_mul(t0, im[4 * r]), d_mul(t1, im[4 * r + 1])), d_mul(t2, im[4 * r + 2])), im[4 * r + 3]);
        }
        float tmp[3];
        for (int d = 0; d < 3; ++d) {
            const float l = f_sub(loc[d], (float)mn[d]);
            const float a = f_sub(l, border[d]);
            const float b = f_sub(f_sub((float)dim_minus1[d], l), border[d]);
            tmp[d] = a < b ? a : b;
            if (tmp[d] <= 0.f) { out[i] = 0.f; return; }
        }
        float min_distance = 1.f;
        for (int d = 0; d < 3; ++d) {
            const float rel = f_div(tmp[d], blending[d]);
            if (rel < 1.f) min_distance = (float)((double)min_distance * lut[(int)((double)rel * 1000.0 + 0.5)]);
        }
        out[i] = min_distance;
    }
};
void blend_weights(stream_t s, float* out, const double* lut_dev, const int vol[3], const int goff[3], const int box_min[3], const int box_max[3],
                   const float border[3], const float blending[3], const double* inv_affine, const int* bbox_offset) {
    BlendKernel k;
    k.out = out; k.lut = lut_dev; k.nx = vol[0]; k.ny = vol[1];
    k.affine = inv_affine ? 1 : 0;
    for (int i = 0; i < 12; ++i) k.im[i] = inv_affine ? inv_affine[i] : 0.0;
    for (int d = 0; d < 3; ++d) {
        k.offset[d] = bbox_offset ? bbox_offset[d] : 0;
        k.goff[d] = goff[d]; k.mn[d] = box_min[d]; k.dim_minus1[d] = box_max[d] - box_min[d];
        k.border[d] = border[d]; k.blending[d] = blending[d];
    }
    pfor((long long)vol[0] * vol[1] * vol[2], k, s);
}
std::vector<double> blend_lut() {      // BlendingRealRandomAccess.java:47-56, including the accumulating loop variable
    std::vector<double> lut(1001, 0.0);
    for (double d = 0; d <= 1.0001; d = d + 0.001) {
        const int idx = (int)(d * 1000.0 + 0.5);
        if (idx <= 1000) lut[idx] = (std::cos((1 - d) * M_PI) + 1) / 2;
    }
    return lut;
}

// ---------------------------------------------------------------------------------------------------------------------
// View materialisation: ProcessInputImages.fuseGroups for one group on the device
// (M/process/deconvolution/util/ProcessInputImages.java:307-393).  Per fused voxel and raw view: world = voxel + bbox min,
// t = inverse affine (double, left to right; TransformedInputRandomAccess.java:63-69); image sample = n-linear / nearest of
// the raw view where t is strictly inside (AbstractTransformedIntervalRandomAccess.java:73-84), max(minValue, .)
// (AbstractTransformedImgRandomAccess.java:76-91), else the outside value; FusedRandomAccess AVG with the fusion blending
// weights (FusedRandomAccess.java:66-91) -> image, CombineWeightsSumRandomAccess of the deconvolution blending weights
// (weightcombination/CombineWeightsSumRandomAccess.java:40-51) -> weight.  The samplers restate imglib2 8.0.0's
// NLinearInterpolator3D / NearestNeighborInterpolator for FloatType (see oracle/mvdecon_oracle.py, same section).
// ---------------------------------------------------------------------------------------------------------------------
MVD_HD float blend_weight_at(const float loc[3], const int dim_minus1[3], const float border[3], const float blending[3], const double* lut) {
    float tmp[3];
    for (int d = 0; d < 3; ++d) {                     // image interval min is 0 (views from the ImgLoader are zero-min)
        const float l = f_sub(loc[d], 0.f);
        const float a = f_sub(l, border[d]);
        const float b = f_sub(f_sub((float)dim_minus1[d], l), border[d]);
        tmp[d] = a < b ? a : b;
        if (tmp[d] <= 0.f) return 0.f;
    }
    float min_distance = 1.f;
    for (int d = 0; d < 3; ++d) {
        const float rel = f_div(tmp[d], blending[d]);
        if (rel < 1.f) min_distance = (float)((double)min_distance * lut[(int)((double)rel * 1000.0 + 0.5)]);
    }
    return min_distance;
}

MVD_HD float nlinear3_at(const float* raw, const int dims[3], double t0, double t1, double t2) {
    const double f0 = floor(t0), f1 = floor(t1), f2 = floor(t2);
    const double w0 = t0 - f0, w1 = t1 - f1, w2 = t2 - f2;
    const double w0i = 1.0 - w0, w1i = 1.0 - w1, w2i = 1.0 - w2;
    const long long x = (long long)f0, y = (long long)f1, z = (long long)f2;
    const long long sy = dims[0], sz = (long long)dims[0] * dims[1];
    const float* p = raw + z * sz + y * sy + x;
    // Gray-code corner order of NLinearInterpolator3D; every product is rounded to float (FloatType.mul(double)) before the float add
    float acc = (float)d_mul((double)p[0], d_mul(d_mul(w0i, w1i), w2i));
    acc = f_add(acc, (float)d_mul((double)p[1], d_mul(d_mul(w0, w1i), w2i)));
    acc = f_add(acc, (float)d_mul((double)p[1 + sy], d_mul(d_mul(w0, w1), w2i)));
    acc = f_add(acc, (float)d_mul((double)p[sy], d_mul(d_mul(w0i, w1), w2i)));
    acc = f_add(acc, (float)d_mul((double)p[sy + sz], d_mul(d_mul(w0i, w1), w2)));
    acc = f_add(acc, (float)d_mul((double)p[1 + sy + sz], d_mul(d_mul(w0, w1), w2)));
    acc = f_add(acc, (float)d_mul((double)p[1 + sz], d_mul(d_mul(w0, w1i), w2)));
    acc = f_add(acc, (float)d_mul((double)p[sz], d_mul(d_mul(w0i, w1i), w2)));
    return acc;
}

struct FuseGroupKernel {
    RawViewDev inl[4];                 // the first views travel as kernel parameters (constant bank), the rest is read from global memory
    const RawViewDev* views; int count;
    float* img_out; float* w_out; const double* lut;
    int nx, ny; int goff[3]; int bbox_min[3];
    float min_value, outside_value;
    MVD_HD void one(const RawViewDev& v, double s0, double s1, double s2, double& sum_i, double& sum_w, double& sum_d) const {
        double t[3];
        for (int r = 0; r < 3; ++r)
            t[r] = d_add(d_add(d_add(d_mul(s0, v.im[4 * r]), d_mul(s1, v.im[4 * r + 1])), d_mul(s2, v.im[4 * r + 2])), v.im[4 * r + 3]);
        float val = outside_value;
        if (t[0] > 0 && t[1] > 0 && t[2] > 0 && t[0] < (double)(v.dims[0] - 1) && t[1] < (double)(v.dims[1] - 1) && t[2] < (double)(v.dims[2] - 1)) {
            float smp;
            if (v.interpolation == 1) smp = nlinear3_at(v.raw, v.dims, t[0], t[1], t[2]);
            else {                                  // Util.roundToLong: half away from zero (t > 0 here)
                const long long rx = (long long)(t[0] + 0.5), ry = (long long)(t[1] + 0.5), rz = (long long)(t[2] + 0.5);
                smp = v.raw[(rz * v.dims[1] + ry) * v.dims[0] + rx];
            }
            val = smp > min_value ? smp : min_value;                                   // Math.max(minValue, sample)
        }
        const float loc[3] = {(float)t[0], (float)t[1], (float)t[2]};                     // TransformedRasteredRandomAccess.java:96-114
        const int dm1[3] = {v.dims[0] - 1, v.dims[1] - 1, v.dims[2] - 1};
        const float wf = v.fusion_blend ? blend_weight_at(loc, dm1, v.fusion_border, v.fusion_range, lut) : 1.f;
        const float wd = v.decon_blend ? blend_weight_at(loc, dm1, v.decon_border, v.decon_range, lut) : 1.f;
        if (wf != 0.f) { sum_i = d_add(sum_i, d_mul((double)val, (double)wf)); sum_w = d_add(sum_w, (double)wf); }
        sum_d = d_add(sum_d, (double)wd);
    }
    MVD_HD void operator()(long long i) const {
        int x, y, z;
        if (i < 0xFFFFFFFFLL) {                         // 32-bit index arithmetic whenever the box has fewer than 2^32 voxels
            const unsigned u = (unsigned)i, r = u / (unsigned)nx;
            x = (int)(u - r * (unsigned)nx); z = (int)(r / (unsigned)ny); y = (int)(r - (unsigned)z * (unsigned)ny);
        } else {
            x = (int)(i % nx); y = (int)((i / nx) % ny); z = (int)(i / ((long long)nx * ny));
        }
        const double s0 = (double)((long long)x + goff[0] + bbox_min[0]), s1 = (double)((long long)y + goff[1] + bbox_min[1]),
                     s2 = (double)((long long)z + goff[2] + bbox_min[2]);
        double sum_i = 0, sum_w = 0, sum_d = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = 0; j < 4; ++j)
            if (j < count) one(inl[j], s0, s1, s2, sum_i, sum_w, sum_d);
        for (int j = 4; j < count; ++j) one(views[j], s0, s1, s2, sum_i, sum_w, sum_d);
        img_out[i] = sum_w > 0 ? (float)(sum_i / sum_w) : 0.f;
        w_out[i] = (float)sum_d;
    }
};
void fuse_group(stream_t s, const RawViewDev* views_dev, const RawViewDev* views_host, int count, float* img_out, float* w_out,
                const double* lut_dev, const int vol[3], const int goff[3], const int bbox_min[3], float min_value, float outside_value) {
    FuseGroupKernel k;
    std::memset(&k, 0, sizeof(k));
    for (int j = 0; j < count && j < 4; ++j) k.inl[j] = views_host[j];     // same records (device raw pointers) as views_dev
    k.views = views_dev; k.count = count; k.img_out = img_out; k.w_out = w_out; k.lut = lut_dev;
    k.nx = vol[0]; k.ny = vol[1];
    for (int d = 0; d < 3; ++d) { k.goff[d] = goff[d]; k.bbox_min[d] = bbox_min[d]; }
    k.min_value = min_value; k.outside_value = outside_value;
    pfor((long long)vol[0] * vol[1] * vol[2], k, s);
}

// NormalizingRandomAccess.get for every view, in place (normalization/NormalizingRandomAccess.java:75-109,183-214)
struct NormalizeWeights {
    WeightPtrs w; int V; double osem; int smooth; float max_diff_range, scaling_range;
    MVD_HD void operator()(long long i) const {
        double sumW = 0;
        float u[MVD_MAX_VIEWS];
        for (int j = 0; j < V; ++j) {
            const double value = (double)w.w[j][i] < 1.0 ? (double)w.w[j][i] : 1.0;      // Math.min(1.0, raw)
            u[j] = (float)value;
            sumW += value;
        }
        for (int j = 0; j < V; ++j) {
            double v;
            if (smooth) {
                if (sumW <= 0) v = 0;
                else {
                    const float ideal = (float)((double)u[j] / sumW);
                    const float diff = f_sub(u[j], ideal);
                    const float ad = diff < 0.f ? -diff : diff;
                    float y = f_mul(f_sub(max_diff_range, ad), f_div(1.0f, max_diff_range));
                    y = y > 0.f ? y : 0.f;
                    const float scale = f_mul(f_mul(y, u[j]), scaling_range);
                    v = (double)f_sub(u[j] < ideal ? u[j] : ideal, scale);
                }
            } else if (sumW > 1) v = (double)(float)((double)u[j] / sumW);                  // hardWeights returns float
            else v = (double)u[j];
            const double r = v * osem;
            w.w[j][i] = (float)(r < 1.0 ? r : 1.0);
        }
    }
};
void normalize_weights(stream_t s, const WeightPtrs& w, int V, long long n, double osem, bool smooth, float mdr, float sr) {
    pfor(n, NormalizeWeights{w, V, osem, smooth ? 1 : 0, mdr, sr}, s);
}

// ---------------------------------------------------------------------------------------------------------------------
// computeNextValueMul (M/process/deconvolution/iteration/mul/DeconvolutionMethods.java:370-419) + statistics
// ---------------------------------------------------------------------------------------------------------------------
MVD_HD float next_psi_value_mul(float last, const MulPtrs& p, int V, long long i, float lambda, float min_value, float max_intensity) {
    double sumW = 0, prod = 1;
    for (int j = 0; j < V; ++j) { prod *= (double)p.integral[j][i]; sumW += (double)p.weight[j][i]; }
#if defined(__CUDA_ARCH__)
    prod = pow(prod, 1.0 / (double)V);
#else
    prod = std::pow(prod, 1.0 / (double)V);
#endif
    sumW = sumW < 1.0 ? sumW : 1.0;
    const float value = f_mul(last, (float)prod);
    float adjusted;
    if (value > 0.f) {
        if (lambda > 0.f) adjusted = f_mul((float)d_tikhonov((double)f_div(value, max_intensity), (double)lambda), max_intensity);
        else adjusted = value;
    } else adjusted = min_value;
    float nxt;
    if (f_isnan(adjusted)) nxt = min_value;
    else nxt = (min_value > adjusted) ? min_value : adjusted;
    return f_add(last, f_mul(f_sub(nxt, last), (float)sumW));
}
struct FinishMulStats {
    double* s; const float* m;
    MVD_HD void operator()(long long) const { s[1] = (double)*m; }
};
#ifndef MVD_HOST_EMU
__global__ void mul_combine_kernel(MulPtrs p, int V, const float* psi_in, float* psi_out, long long n, long long own0, long long own1,
                                   float lambda, float min_value, float max_intensity, double* stats, float* smax) {
    double s = 0; float m = -1.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float last = psi_in[i];
        if (i < own0 || i >= own1) { psi_out[i] = last; continue; }
        const float nxt = next_psi_value_mul(last, p, V, i, lambda, min_value, max_intensity);
        psi_out[i] = nxt;
        const float change = f_sub(nxt, last);
        s += (double)change; m = fmaxf(m, change);
    }
    block_accumulate(s, 0, m, stats, nullptr, smax);
}
#endif
void mul_combine(stream_t st, const MulPtrs& p, int V, const float* psi_in, float* psi_out, long long n, long long own0, long long own1, float lambda,
                 float min_value, float max_intensity, double* stats_dev /*[2] sum,max*/, float* scratch_max_dev) {
    dev::zero(stats_dev, sizeof(double) * 2, st);
#ifndef MVD_HOST_EMU
    const float init = -1.f;
    MVD_CUDA_CHECK(cudaMemcpyAsync(scratch_max_dev, &init, sizeof(float), cudaMemcpyHostToDevice, st));
    mul_combine_kernel<<<148 * 8, 256, 0, st>>>(p, V, psi_in, psi_out, n, own0, own1, lambda, min_value, max_intensity, stats_dev, scratch_max_dev);
    MVD_CUDA_CHECK(cudaGetLastError());
    pfor(1, FinishMulStats{stats_dev, scratch_max_dev}, st);
#else
    double s = 0; float m = -1.f;
    for (long long i = 0; i < n; ++i) {
        const float last = psi_in[i];
        if (i < own0 || i >= own1) { psi_out[i] = last; continue; }
        const float nxt = next_psi_value_mul(last, p, V, i, lambda, min_value, max_intensity);
        psi_out[i] = nxt;
        const float change = nxt - last;
        s += (double)change; m = std::max(m, change);
    }
    stats_dev[0] = s; stats_dev[1] = (double)m; (void)scratch_max_dev;
#endif
}

// Gauss3 half kernel (ASSUMPTION, third-party net.imglib2.algorithm.gauss3.Gauss3, see oracle/mvdecon_oracle.py:gauss3_halfkernel)
std::vector<double> gauss3_halfkernel(double sigma) {
    const int size = std::max(2, (int)(3 * sigma + 0.5) + 1);
    std::vector<double> k(size);
    k[0] = 1;
    for (int x = 1; x < size; ++x) k[x] = std::exp(-(double)(x * x) / (2 * sigma * sigma));
    if (size > 3) {
        double sqrt_slope = 1e300;
        int r = size;
        while (r > size / 2) {
            --r;
            const double a = std::sqrt(k[r]) / (size - r);
            if (a < sqrt_slope) sqrt_slope = a; else break;
        }
        for (int r1 = r + 2; r1 < size; ++r1) k[r1] = (double)(size - r1) * (size - r1) * sqrt_slope * sqrt_slope;
    }
    double s = 0.5 * k[0];
    for (int x = 1; x < size; ++x) s += k[x];
    s *= 2;
    for (double& v : k) v /= s;
    return k;
}

}  // namespace mvd
