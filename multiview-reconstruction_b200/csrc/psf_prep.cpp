// PSF preparation ahead of the loop (host; a PSF is a few thousand voxels): PSFPreparation.loadGroupTransformPSFs
// (M/process/deconvolution/util/PSFPreparation.java:41-89) = per view normalize + transformPSF, per group computeAverageImage over the
// minimal size, optionally makeSameSize over all groups.  imglib2's samplers / estimateBounds are restated from their published
// algorithms (imglib2 8.0.0, imglib2-realtransform; absent from the reference tree) -- see oracle/mvdecon_oracle.py, same section.
#include <algorithm>
#include <cmath>

#include "engine.h"

namespace mvd {

namespace {
inline void apply_affine(const double m[12], double x, double y, double z, double out[3]) {
    for (int r = 0; r < 3; ++r) {
        volatile double a = x * m[4 * r];
        volatile double b = y * m[4 * r + 1];
        volatile double c = z * m[4 * r + 2];
        volatile double ab = a + b;
        volatile double abc = ab + c;
        out[r] = abc + m[4 * r + 3];
    }
}
inline float sample_zero_ext(const float* p, const int d[3], long long x, long long y, long long z) {
    if (x < 0 || y < 0 || z < 0 || x >= d[0] || y >= d[1] || z >= d[2]) return 0.f;
    return p[(z * d[1] + y) * d[0] + x];
}
inline float mulf(float v, double w) { volatile double r = (double)v * w; return (float)r; }       // FloatType.mul(double)
inline float addf(float a, float b) { volatile float r = a + b; return r; }
}  // namespace

// PSFExtraction.transformPSF (M/process/psf/PSFExtraction.java:367-409): bounds of the transformed interval [0, dim-1] (8 corners),
// newSize = (int)size + 1 made odd, offset = A(dim / 2) - newSize / 2
void psf_transformed_geometry(const int dims[3], const double affine[12], int new_dims[3], double offset[3]) {
    double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
    for (int cz = 0; cz < 2; ++cz)
        for (int cy = 0; cy < 2; ++cy)
            for (int cx = 0; cx < 2; ++cx) {
                double t[3];
                apply_affine(affine, cx ? dims[0] - 1 : 0, cy ? dims[1] - 1 : 0, cz ? dims[2] - 1 : 0, t);
                for (int d = 0; d < 3; ++d) { mn[d] = std::min(mn[d], t[d]); mx[d] = std::max(mx[d], t[d]); }
            }
    double ctr[3];
    apply_affine(affine, dims[0] / 2, dims[1] / 2, dims[2] / 2, ctr);
    for (int d = 0; d < 3; ++d) {
        const double size = mx[d] - mn[d];
        int n = (int)size + 1;
        if (n % 2 == 0) ++n;
        new_dims[d] = n;
        offset[d] = ctr[d] - (double)(n / 2);
    }
}

// PSFExtraction.getTransformedNormalizedPSF (:182-193): normalize (:453-473) then transform (:411-451)
std::vector<float> psf_transform_normalized(const float* psf, const int dims[3], const double affine[12], const double inv_affine[12], int new_dims[3]) {
    const size_t n = (size_t)dims[0] * dims[1] * dims[2];
    double lo = 1.7976931348623157e308, hi = -1.7976931348623157e308;
    for (size_t i = 0; i < n; ++i) { const double v = psf[i]; if (v < lo) lo = v; if (v > hi) hi = v; }
    std::vector<float> p(n);
    for (size_t i = 0; i < n; ++i) p[i] = (float)(((double)psf[i] - lo) / (hi - lo));
    double off[3];
    psf_transformed_geometry(dims, affine, new_dims, off);
    std::vector<float> out((size_t)new_dims[0] * new_dims[1] * new_dims[2]);
    for (int z = 0; z < new_dims[2]; ++z)
        for (int y = 0; y < new_dims[1]; ++y)
            for (int x = 0; x < new_dims[0]; ++x) {
                double t[3];
                apply_affine(inv_affine, (double)x + off[0], (double)y + off[1], (double)z + off[2], t);
                const double f0 = std::floor(t[0]), f1 = std::floor(t[1]), f2 = std::floor(t[2]);
                const double w0 = t[0] - f0, w1 = t[1] - f1, w2 = t[2] - f2, w0i = 1.0 - w0, w1i = 1.0 - w1, w2i = 1.0 - w2;
                const long long X = (long long)f0, Y = (long long)f1, Z = (long long)f2;
                auto S = [&](int dx, int dy, int dz) { return sample_zero_ext(p.data(), dims, X + dx, Y + dy, Z + dz); };
                float acc = mulf(S(0, 0, 0), w0i * w1i * w2i);
                acc = addf(acc, mulf(S(1, 0, 0), w0 * w1i * w2i));
                acc = addf(acc, mulf(S(1, 1, 0), w0 * w1 * w2i));
                acc = addf(acc, mulf(S(0, 1, 0), w0i * w1 * w2i));
                acc = addf(acc, mulf(S(0, 1, 1), w0i * w1 * w2));
                acc = addf(acc, mulf(S(1, 1, 1), w0 * w1 * w2));
                acc = addf(acc, mulf(S(1, 0, 1), w0 * w1i * w2));
                acc = addf(acc, mulf(S(0, 0, 1), w0i * w1i * w2));
                out[((size_t)z * new_dims[1] + y) * new_dims[0] + x] = acc;
            }
    return out;
}

// PSFCombination.computeAverageImage (M/process/psf/PSFCombination.java:74-135)
std::vector<float> psf_average(const float* const* psfs, const int (*dims)[3], int count, bool use_max, int out_dims[3]) {
    if (count < 1) throw Error("no PSFs to average");
    for (int d = 0; d < 3; ++d) {
        out_dims[d] = dims[0][d];
        for (int j = 1; j < count; ++j) out_dims[d] = use_max ? std::max(out_dims[d], dims[j][d]) : std::min(out_dims[d], dims[j][d]);
    }
    std::vector<float> avg((size_t)out_dims[0] * out_dims[1] * out_dims[2], 0.f);
    for (int j = 0; j < count; ++j) {
        const int* pd = dims[j];
        for (int z = 0; z < pd[2]; ++z)
            for (int y = 0; y < pd[1]; ++y)
                for (int x = 0; x < pd[0]; ++x) {
                    const int ax = out_dims[0] / 2 - (pd[0] / 2 - x), ay = out_dims[1] / 2 - (pd[1] / 2 - y), az = out_dims[2] / 2 - (pd[2] / 2 - z);
                    if (ax < 0 || ay < 0 || az < 0 || ax >= out_dims[0] || ay >= out_dims[1] || az >= out_dims[2]) continue;   // extendZero target
                    float& a = avg[((size_t)az * out_dims[1] + ay) * out_dims[0] + ax];
                    a = addf(a, psfs[j][((size_t)z * pd[1] + y) * pd[0] + x]);
                }
    }
    for (float& a : avg) a = (float)((double)a / (double)count);
    return avg;
}

// PSFCombination.makeSameSize (:182-212)
std::vector<float> psf_make_same_size(const float* psf, const int dims[3], const int new_dims[3]) {
    const size_t n = (size_t)dims[0] * dims[1] * dims[2];
    double mn = 1.7976931348623157e308;
    for (size_t i = 0; i < n; ++i) mn = std::min(mn, (double)psf[i]);
    std::vector<float> out((size_t)new_dims[0] * new_dims[1] * new_dims[2]);
    for (int z = 0; z < new_dims[2]; ++z)
        for (int y = 0; y < new_dims[1]; ++y)
            for (int x = 0; x < new_dims[0]; ++x) {
                const int sx = x - new_dims[0] / 2 + dims[0] / 2, sy = y - new_dims[1] / 2 + dims[1] / 2, sz = z - new_dims[2] / 2 + dims[2] / 2;
                const bool in = sx >= 0 && sy >= 0 && sz >= 0 && sx < dims[0] && sy < dims[1] && sz < dims[2];
                out[((size_t)z * new_dims[1] + y) * new_dims[0] + x] = in ? psf[((size_t)sz * dims[1] + sy) * dims[0] + sx] : (float)mn;
            }
    return out;
}

}  // namespace mvd
