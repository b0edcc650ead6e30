// extern "C" surface declared in include/mvdecon.h
#include "../../include/mvdecon.h"

#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>

#include "engine.h"

#include <algorithm>
#include <thread>

using namespace mvd;

struct mvd_comm { std::shared_ptr<NcclComm> comm; };
struct mvd_context {
    Engine* engine = nullptr;
    int halo_lo = 0, halo_hi = 0;
};

namespace {
thread_local std::string g_last_error;

template <class F>
int guarded(F&& f) {
    try {
        f();
        g_last_error.clear();
        return 0;
    } catch (const std::exception& e) {
        g_last_error = e.what();
        return 1;
    } catch (...) {
        g_last_error = "unknown error";
        return 1;
    }
}
void require(bool c, const char* msg) { if (!c) throw Error(msg); }

#ifndef MVD_HOST_EMU
void require_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) throw Error(std::string("no usable CUDA device: ") + cudaGetErrorString(e));
    if (device < 0 || device >= n) throw Error("CUDA device ordinal out of range");
}
#else
void require_device(int) {}
#endif
}  // namespace

extern "C" {

const char* mvd_last_error(void) { return g_last_error.c_str(); }
int mvd_version(void) { return 200; }
int mvd_reference_threads(void) {        // Threads.numThreads() = max(4, Prefs.getThreads()), ImageJ's default = processors (Threads.java:40)
    const unsigned n = std::thread::hardware_concurrency();
    return (int)std::max(4u, n);
}
int mvd_supported_fft_lengths(int* out, int cap) {
    const std::vector<int>& v = supported_lengths();
    for (int i = 0; i < (int)v.size() && i < cap && out; ++i) out[i] = v[i];
    return (int)v.size();
}

int mvd_create(const mvd_config* cfg, mvd_context** out) {
    return guarded([&] {
        require(cfg && out, "null argument");
        require_device(cfg->device);
        Engine::Config c;
        c.device = cfg->device;
        c.num_views = cfg->num_views;
        c.psf_type = cfg->psf_type;
        c.lambda = cfg->lambda;
        c.min_value = cfg->min_value;
        c.max_len = cfg->max_fft_len > 0 ? cfg->max_fft_len : 1152;
        // 0 = what the reference does on this host (Threads.numThreads()), T > 0 = a reference run with T threads, < 0 = exact sums
        c.norm_quirk_threads = cfg->norm_quirk_threads > 0 ? cfg->norm_quirk_threads : (cfg->norm_quirk_threads < 0 ? 0 : mvd_reference_threads());
        require(cfg->exchange_scheme == 0 || cfg->exchange_scheme == 1, "exchange_scheme must be 0 or 1");
        c.exchange_scheme = cfg->exchange_scheme;
        require(cfg->psf_type >= 0 && cfg->psf_type <= 3, "bad psf_type");
        Geometry& g = c.geom;
        for (int d = 0; d < 3; ++d) {
            g.gdim[d] = cfg->dims[d]; g.goff[d] = 0; g.vol[d] = cfg->dims[d]; g.own_lo[d] = 0; g.own_hi[d] = cfg->dims[d];
        }
        int slo = cfg->shard_lo, shi = cfg->shard_hi, z0 = cfg->local_z0, nz = cfg->local_nz;
        if (shi <= slo) { slo = 0; shi = cfg->dims[2]; }
        if (nz <= 0) { z0 = 0; nz = cfg->dims[2]; }
        g.own_lo[2] = slo; g.own_hi[2] = shi; g.goff[2] = z0; g.vol[2] = nz;
        if (cfg->shard_y_hi > cfg->shard_y_lo && cfg->local_ny > 0) {
            g.own_lo[1] = cfg->shard_y_lo; g.own_hi[1] = cfg->shard_y_hi; g.goff[1] = cfg->local_y0; g.vol[1] = cfg->local_ny;
        }
        mvd_context* ctx = new mvd_context();
        try { ctx->engine = new Engine(c); } catch (...) { delete ctx; throw; }
        *out = ctx;
    });
}

int mvd_destroy(mvd_context* ctx) {
    return guarded([&] {
        if (!ctx) return;
        delete ctx->engine;
        delete ctx;
    });
}

int mvd_set_view(mvd_context* ctx, int v, const float* img, const float* weight) {
    return guarded([&] { require(ctx && img, "null argument"); ctx->engine->set_view_host(v, img, weight); });
}
int mvd_set_view_async(mvd_context* ctx, int v, const float* img, const float* weight) {
    return guarded([&] { require(ctx && img, "null argument"); ctx->engine->set_view_host_async(v, img, weight); });
}
int mvd_set_view_device(mvd_context* ctx, int v, const float* img, const float* weight) {
    return guarded([&] { require(ctx && img, "null argument"); ctx->engine->set_view_device(v, img, weight); });
}
int mvd_set_psf(mvd_context* ctx, int v, const float* psf, const int kdims[3]) {
    return guarded([&] {
        require(ctx && psf && kdims, "null argument");
        require(kdims[0] > 0 && kdims[1] > 0 && kdims[2] > 0, "empty PSF");
        ctx->engine->set_psf(v, psf, kdims);
    });
}
int mvd_set_kernels(mvd_context* ctx, int v, const float* k1, const int k1d[3], const float* k2, const int k2d[3]) {
    return guarded([&] { require(ctx && k1 && k2 && k1d && k2d, "null argument"); ctx->engine->set_kernels(v, k1, k1d, k2, k2d); });
}
int mvd_init_views(mvd_context* ctx) {
    return guarded([&] {
        require(ctx, "null context");
        ctx->engine->init_views();
        ctx->engine->halo_needed(ctx->halo_lo, ctx->halo_hi);
    });
}
int mvd_get_kernel_dims(mvd_context* ctx, int v, int which, int kdims[3]) {
    return guarded([&] { require(ctx && kdims && (which == 1 || which == 2), "bad argument"); ctx->engine->get_kernel_dims(v, which, kdims); });
}
int mvd_get_kernel(mvd_context* ctx, int v, int which, float* out) {
    return guarded([&] { require(ctx && out && (which == 1 || which == 2), "bad argument"); ctx->engine->get_kernel(v, which, out); });
}
int mvd_set_psi(mvd_context* ctx, const float* psi) {
    return guarded([&] { require(ctx && psi, "null argument"); ctx->engine->set_psi_host(psi); });
}
int mvd_get_psi(mvd_context* ctx, float* psi) {
    return guarded([&] { require(ctx && psi, "null argument"); ctx->engine->get_psi_host(psi); });
}
int mvd_set_max_intensities(mvd_context* ctx, const float* mx) {
    return guarded([&] { require(ctx && mx, "null argument"); ctx->engine->set_max_intensity(mx); });
}
int mvd_psi_init(mvd_context* ctx, int type, double sigma, double* avg_out, float* max_out) {
    return guarded([&] { require(ctx, "null context"); ctx->engine->psi_init(type, sigma, avg_out, max_out); });
}
int mvd_make_blending_weights(mvd_context* ctx, int v, const int box_min[3], const int box_max[3], const float border[3], const float blending[3]) {
    return guarded([&] {
        require(ctx && box_min && box_max && border && blending, "null argument");
        for (int d = 0; d < 3; ++d) require(blending[d] > 0.f, "blending range must be positive");
        ctx->engine->make_blending_weights(v, box_min, box_max, border, blending);
    });
}
int mvd_make_blending_weights_affine(mvd_context* ctx, int v, const int img_min[3], const int img_max[3], const float border[3],
                                     const float blending[3], const double inv_affine[12], const int bbox_offset[3]) {
    return guarded([&] {
        require(ctx && img_min && img_max && border && blending && inv_affine && bbox_offset, "null argument");
        for (int d = 0; d < 3; ++d) require(blending[d] > 0.f, "blending range must be positive");
        ctx->engine->make_blending_weights(v, img_min, img_max, border, blending, inv_affine, bbox_offset);
    });
}
int mvd_normalize_weights(mvd_context* ctx, double osem_speedup, int additional_smooth, float max_diff_range, float scaling_range) {
    return guarded([&] { require(ctx, "null context"); ctx->engine->normalize_view_weights(osem_speedup, additional_smooth != 0, max_diff_range, scaling_range); });
}
int mvd_get_weight(mvd_context* ctx, int v, float* out) {
    return guarded([&] { require(ctx && out, "null argument"); ctx->engine->get_weight_host(v, out); });
}
int mvd_get_image(mvd_context* ctx, int v, float* out) {
    return guarded([&] { require(ctx && out, "null argument"); ctx->engine->get_image_host(v, out); });
}
int mvd_run_iteration_mul(mvd_context* ctx, double stats[2]) {
    return guarded([&] {
        require(ctx, "null context");
        ctx->engine->iteration_mul();
        IterStats s{0, -1};
        ctx->engine->fetch_stats(1, &s);
        if (stats) { stats[0] = s.sum_change; stats[1] = s.max_change; }
    });
}
int mvd_run_view_update(mvd_context* ctx, int v, double stats[2]) {
    return guarded([&] {
        require(ctx, "null context");
        ctx->engine->view_update(v);
        IterStats s{0, -1};
        ctx->engine->fetch_stats(1, &s);
        if (stats) { stats[0] = s.sum_change; stats[1] = s.max_change; }
    });
}
int mvd_skip_empty_tiles(mvd_context* ctx, int on, int* skipped_out) {
    return guarded([&] {
        require(ctx, "null context");
        const int n = ctx->engine->skip_empty_tiles(on != 0);
        if (skipped_out) *skipped_out = n;
    });
}
int mvd_run_iterations(mvd_context* ctx, int n, double* stats) {
    return guarded([&] {
        require(ctx && n >= 0, "bad argument");
        std::vector<IterStats> s((size_t)n * ctx->engine->num_views() + 1);
        ctx->engine->run_iterations(n, stats ? s.data() : nullptr);
        if (stats)
            for (size_t i = 0; i < (size_t)n * ctx->engine->num_views(); ++i) { stats[2 * i] = s[i].sum_change; stats[2 * i + 1] = s[i].max_change; }
    });
}
int mvd_enqueue_view_update(mvd_context* ctx, int v) {
    return guarded([&] { require(ctx, "null context"); ctx->engine->view_update(v); });
}
int mvd_synchronize(mvd_context* ctx) {
    return guarded([&] { require(ctx, "null context"); ctx->engine->synchronize(); });
}
int mvd_fetch_stats(mvd_context* ctx, int count, double* stats) {
    return guarded([&] {
        require(ctx && stats && count >= 0, "bad argument");
        std::vector<IterStats> s((size_t)count + 1, IterStats{0, -1});
        ctx->engine->fetch_stats(count, s.data());
        for (int i = 0; i < count; ++i) { stats[2 * i] = s[i].sum_change; stats[2 * i + 1] = s[i].max_change; }
    });
}
int mvd_tile_info(mvd_context* ctx, int tile_dims[3], int* num_tiles, double* ratio, int* launches) {
    return guarded([&] {
        require(ctx && ctx->engine->convolver(), "views not initialised");
        const Convolver* c = ctx->engine->convolver();
        if (tile_dims) for (int d = 0; d < 3; ++d) tile_dims[d] = c->tile_dims()[d];
        if (num_tiles) *num_tiles = c->num_tiles();
        if (ratio) *ratio = c->fft_volume_ratio();
        if (launches) *launches = ctx->engine->launches_per_view_update();
    });
}
int mvd_halo_planes(mvd_context* ctx, int* lo, int* hi) {
    return guarded([&] {
        require(ctx && ctx->engine->convolver(), "views not initialised");
        if (lo) *lo = ctx->halo_lo;
        if (hi) *hi = ctx->halo_hi;
    });
}
int mvd_halo_rows(mvd_context* ctx, int* lo, int* hi) {
    return guarded([&] {
        require(ctx && ctx->engine->convolver(), "views not initialised");
        int l, h;
        ctx->engine->halo_needed_y(l, h);
        if (lo) *lo = l;
        if (hi) *hi = h;
    });
}
int mvd_psi_device_ptr(mvd_context* ctx, void** current) {
    return guarded([&] { require(ctx && current, "null argument"); *current = ctx->engine->psi_device(); });
}
int mvd_stream_handle(mvd_context* ctx, void** s) {
    return guarded([&] { require(ctx && s, "null argument"); *s = (void*)ctx->engine->stream(); });
}

int mvd_comm_unique_id(char id_out[128]) {
    return guarded([&] { require(id_out != nullptr, "null argument"); NcclComm::unique_id(id_out); });
}
int mvd_comm_create(const char id[128], int world, int rank, int device, mvd_comm** out) {
    return guarded([&] {
        require(id && out, "null argument");
        require_device(device);
        mvd_comm* c = new mvd_comm();
        try { c->comm = std::make_shared<NcclComm>(id, world, rank, device); } catch (...) { delete c; throw; }
        *out = c;
    });
}
int mvd_comm_destroy(mvd_comm* comm) {
    return guarded([&] { delete comm; });
}
int mvd_comm_attach(mvd_context* ctx, mvd_comm* comm, int py, int pz) {
    return guarded([&] { require(ctx && comm, "null argument"); ctx->engine->comm_attach(comm->comm, py, pz); });
}
static_assert(sizeof(mvd_halo_box) == sizeof(HaloBox), "mvd_halo_box mirrors HaloBox");
int mvd_set_exchange_callback(mvd_context* ctx, mvd_exchange_fn fn, void* user) {
    return guarded([&] {
        require(ctx != nullptr, "null argument");
        ctx->engine->set_exchange_callback(reinterpret_cast<ExchangeFn>(fn), user);
    });
}
int mvd_set_reduce_callback(mvd_context* ctx, mvd_reduce_fn fn, void* user) {
    return guarded([&] { require(ctx != nullptr, "null argument"); ctx->engine->set_reduce_callback(reinterpret_cast<ReduceFn>(fn), user); });
}
static_assert(sizeof(mvd_raw_view) == sizeof(RawViewDev), "mvd_raw_view mirrors RawViewDev");
int mvd_fuse_group(mvd_context* ctx, int v, const mvd_raw_view* views, int count, const int bbox_min[3], float min_value_img,
                   float outside_value) {
    return guarded([&] {
        require(ctx && views && bbox_min, "null argument");
        ctx->engine->fuse_group_host(v, reinterpret_cast<const RawViewDev*>(views), count, bbox_min, min_value_img, outside_value);
    });
}
int mvd_last_fuse_group_ms(mvd_context* ctx, double* ms) {
    return guarded([&] { require(ctx && ms, "null argument"); *ms = ctx->engine->last_fuse_group_ms(); });
}
int mvd_psf_transformed_dims(const int dims[3], const double affine[12], int new_dims[3]) {
    return guarded([&] {
        require(dims && affine && new_dims, "null argument");
        double off[3];
        psf_transformed_geometry(dims, affine, new_dims, off);
    });
}
int mvd_psf_transform(const float* psf, const int dims[3], const double affine[12], const double inv_affine[12], float* out) {
    return guarded([&] {
        require(psf && dims && affine && inv_affine && out, "null argument");
        require(dims[0] > 0 && dims[1] > 0 && dims[2] > 0, "empty PSF");
        int nd[3];
        const std::vector<float> r = psf_transform_normalized(psf, dims, affine, inv_affine, nd);
        std::copy(r.begin(), r.end(), out);
    });
}
int mvd_psf_average(const float* const* psfs, const int* dims, int count, int use_max, int out_dims[3], float* out) {
    return guarded([&] {
        require(psfs && dims && out_dims && count > 0, "null argument");
        if (!out) {
            for (int d = 0; d < 3; ++d) {
                out_dims[d] = dims[d];
                for (int j = 1; j < count; ++j) out_dims[d] = use_max ? std::max(out_dims[d], dims[3 * j + d]) : std::min(out_dims[d], dims[3 * j + d]);
            }
            return;
        }
        const std::vector<float> r = psf_average(psfs, reinterpret_cast<const int(*)[3]>(dims), count, use_max != 0, out_dims);
        std::copy(r.begin(), r.end(), out);
    });
}
int mvd_psf_make_same_size(const float* psf, const int dims[3], const int new_dims[3], float* out) {
    return guarded([&] {
        require(psf && dims && new_dims && out, "null argument");
        const std::vector<float> r = psf_make_same_size(psf, dims, new_dims);
        std::copy(r.begin(), r.end(), out);
    });
}
int mvd_psi_init_from_file(mvd_context* ctx, const char* path, int precise, double* avg_out, float* max_out) {
    return guarded([&] {
        require(ctx && path, "null argument");
        Engine& e = *ctx->engine;
        const Geometry& g = e.config().geom;
        int d[3];
        const std::vector<float> img = tiff_read_f32(path, d);
        for (int a = 0; a < 3; ++a) {
            require(g.vol[a] == g.gdim[a], "PsiInitFromFile needs an unsharded context");
            if (d[a] != g.gdim[a]) throw Error("Image dimensions do not match: the start image differs from the deconvolved volume");
        }
        e.set_psi_host(img.data());
        e.psi_init(precise ? PSI_AVG : PSI_APPROX_AVG, 0.0, avg_out, max_out, false);
    });
}
int mvd_tiff_dims(const char* path, int dims[3]) {
    return guarded([&] { require(path && dims, "null argument"); tiff_dims(path, dims); });
}
int mvd_tiff_read(const char* path, float* out) {
    return guarded([&] {
        require(path && out, "null argument");
        int d[3];
        const std::vector<float> img = tiff_read_f32(path, d);
        std::copy(img.begin(), img.end(), out);
    });
}
int mvd_tiff_write(const char* path, const float* data, const int dims[3]) {
    return guarded([&] { require(path && data && dims, "null argument"); tiff_write_f32(path, data, dims); });
}
int mvd_n5_dims(const char* dataset_dir, int dims[3]) {
    return guarded([&] { require(dataset_dir && dims, "null argument"); n5_dims(dataset_dir, dims); });
}
int mvd_n5_read(const char* dataset_dir, float* out) {
    return guarded([&] {
        require(dataset_dir && out, "null argument");
        int d[3];
        const std::vector<float> img = n5_read_f32(dataset_dir, d);
        std::copy(img.begin(), img.end(), out);
    });
}
int mvd_n5_write(const char* dataset_dir, const float* data, const int dims[3], const int block_size[3], int gzip_level) {
    return guarded([&] { require(dataset_dir && data && dims && block_size, "null argument"); n5_write_f32(dataset_dir, data, dims, block_size, gzip_level); });
}
int mvd_zarr_write(const char* path, const float* data, const int dims[3], const int chunk_size[3], int gzip_level, const double voxel_size[3]) {
    return guarded([&] { require(path && data && dims && chunk_size, "null argument"); zarr_write_f32(path, data, dims, chunk_size, gzip_level, voxel_size); });
}
int mvd_plan_axis(int gdim, int own_lo, int own_hi, int r1_lo, int r1_hi, int r2_lo, int r2_hi, int is_x, int max_fft_len, int two_exchanges,
                  int* tile_len, int* tiles, int cap, int* num_tiles) {
    return guarded([&] {
        require(tile_len && num_tiles, "null argument");
        require(gdim > 0 && own_lo >= 0 && own_hi <= gdim && own_lo < own_hi, "bad axis range");
        require(r1_lo >= 0 && r1_hi >= 0 && r2_lo >= 0 && r2_hi >= 0, "negative reach");
        const AxisTiling t = plan_axis(gdim, own_lo, own_hi, Reach{r1_lo, r1_hi}, Reach{r2_lo, r2_hi}, is_x != 0, max_fft_len > 0 ? max_fft_len : 1152,
                                       two_exchanges != 0);
        *tile_len = t.T;
        *num_tiles = (int)t.tiles.size();
        for (int i = 0; tiles && i < cap && i < (int)t.tiles.size(); ++i) {
            tiles[3 * i] = t.tiles[(size_t)i].org; tiles[3 * i + 1] = t.tiles[(size_t)i].lo; tiles[3 * i + 2] = t.tiles[(size_t)i].hi;
        }
    });
}
int mvd_exchange_transport(mvd_context* ctx, int* transport) {
    return guarded([&] { require(ctx && transport, "null argument"); *transport = ctx->engine->exchange_transport(); });
}
int mvd_exchange_halos(mvd_context* ctx) {
    return guarded([&] { require(ctx, "null context"); ctx->engine->exchange_halos(); });
}
int mvd_set_profiling(mvd_context* ctx, int on) {
    return guarded([&] { require(ctx && ctx->engine->convolver(), "views not initialised"); ctx->engine->convolver()->set_profiling(on != 0); });
}
int mvd_get_pass_times(mvd_context* ctx, double ms[9], long long counts[9], int reset) {
    return guarded([&] {
        require(ctx && ms && counts && ctx->engine->convolver(), "bad argument");
        ctx->engine->convolver()->collect_pass_times(ms, counts, reset != 0);
    });
}

int mvd_get_aux_times(mvd_context* ctx, double ms[3], long long counts[3], int reset) {
    return guarded([&] {
        require(ctx && ms && counts && ctx->engine->convolver(), "bad argument");
        ctx->engine->convolver()->collect_aux_times(ms, counts, reset != 0);
    });
}

int mvd_convolve(int device, const float* img, const int dims[3], const float* kernel, const int kdims[3], int ext,
                 float ext_value, float* out) {
    return guarded([&] {
        require(img && dims && kernel && kdims && out, "null argument");
        require(ext >= 0 && ext <= 2, "bad extension mode");
        require_device(device);
        dev::set_device(device);
        stream_t s = dev::stream_create();
        try {
            Tables tables(s);
            convolve_host(device, s, &tables, 1152, img, dims, kernel, kdims, ext, ext_value, out, false);
        } catch (...) { dev::stream_destroy(s); throw; }
        dev::stream_destroy(s);
    });
}

int mvd_block_iteration(int device, float* psi_block, const float* img_block, const float* weight_block, const int bd[3],
                        const float* k1, const int k1d[3], const float* k2, const int k2d[3], float lambda, float min_value,
                        float max_intensity, double stats[2]) {
    return guarded([&] {
        require(psi_block && img_block && weight_block && bd && k1 && k2 && k1d && k2d, "null argument");
        require_device(device);
        Engine::Config c;
        c.device = device;
        c.num_views = 1;
        c.psf_type = INDEPENDENT;
        c.lambda = lambda;
        c.min_value = min_value;
        for (int d = 0; d < 3; ++d) { c.geom.gdim[d] = c.geom.vol[d] = bd[d]; c.geom.goff[d] = 0; c.geom.own_lo[d] = 0; c.geom.own_hi[d] = bd[d]; }
        Engine e(c);
        e.set_view_host(0, img_block, weight_block);
        e.set_kernels(0, k1, k1d, k2, k2d);
        e.init_views();
        e.set_max_intensity(&max_intensity);
        e.set_psi_host(psi_block);
        e.view_update(0);
        IterStats s{0, -1};
        e.fetch_stats(1, &s);
        e.get_psi_host(psi_block);
        if (stats) { stats[0] = s.sum_change; stats[1] = s.max_change; }
    });
}

// ---------------------------------------------------------------------------------------------------------------
// L1 legacy symbols
// ---------------------------------------------------------------------------------------------------------------
static void legacy_convolve(const float* im, const int* imDim, const float* kernel, const int* kernelDim, int devCUDA, float* out) {
    const int dims[3] = {imDim[2], imDim[1], imDim[0]};         // {z,y,x} -> (x,y,z)  (CUDATools.java:41-49)
    const int kd[3] = {kernelDim[2], kernelDim[1], kernelDim[0]};
    require_device(devCUDA);
    dev::set_device(devCUDA);
    stream_t s = dev::stream_create();
    try {
        Tables tables(s);
        convolve_host(devCUDA, s, &tables, 1152, im, dims, kernel, kd, EXT_ZERO, 0.f, out, true);
    } catch (...) { dev::stream_destroy(s); throw; }
    dev::stream_destroy(s);
}

void convolution3DfftCUDAInPlace(float* im, int* imDim, float* kernel, int* kernelDim, int devCUDA) {
    int rc = guarded([&] {
        require(im && imDim && kernel && kernelDim, "null argument");
        legacy_convolve(im, imDim, kernel, kernelDim, devCUDA, im);
    });
    if (rc) std::fprintf(stderr, "convolution3DfftCUDAInPlace failed: %s\n", g_last_error.c_str());
}

float* convolution3DfftCUDA(float* im, int* imDim, float* kernel, int* kernelDim, int devCUDA) {
    float* out = nullptr;
    int rc = guarded([&] {
        require(im && imDim && kernel && kernelDim, "null argument");
        const size_t n = (size_t)imDim[0] * imDim[1] * imDim[2];
        out = (float*)std::malloc(sizeof(float) * n);
        require(out != nullptr, "out of host memory");
        legacy_convolve(im, imDim, kernel, kernelDim, devCUDA, out);
    });
    if (rc) {
        std::fprintf(stderr, "convolution3DfftCUDA failed: %s\n", g_last_error.c_str());
        std::free(out);
        return nullptr;
    }
    return out;
}

#ifndef MVD_HOST_EMU
int getNumDevicesCUDA(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return -1;
    return n;
}
static bool get_props(int dev_, cudaDeviceProp& p) { return cudaGetDeviceProperties(&p, dev_) == cudaSuccess; }
int getCUDAcomputeCapabilityMajorVersion(int devCUDA) { cudaDeviceProp p; return get_props(devCUDA, p) ? p.major : -1; }
int getCUDAcomputeCapabilityMinorVersion(int devCUDA) { cudaDeviceProp p; return get_props(devCUDA, p) ? p.minor : -1; }
void getNameDeviceCUDA(int devCUDA, char* name) {
    if (!name) return;
    cudaDeviceProp p;
    if (!get_props(devCUDA, p)) { name[0] = 0; return; }
    std::strncpy(name, p.name, 255);
    name[255] = 0;
}
long long getMemDeviceCUDA(int devCUDA) { cudaDeviceProp p; return get_props(devCUDA, p) ? (long long)p.totalGlobalMem : -1; }
long long getFreeMemDeviceCUDA(int devCUDA) {
    size_t f = 0, t = 0;
    if (cudaSetDevice(devCUDA) != cudaSuccess || cudaMemGetInfo(&f, &t) != cudaSuccess) return -1;
    return (long long)f;
}
#else
int getNumDevicesCUDA(void) { return 1; }
int getCUDAcomputeCapabilityMajorVersion(int) { return 0; }
int getCUDAcomputeCapabilityMinorVersion(int) { return 0; }
void getNameDeviceCUDA(int, char* name) { if (name) std::strcpy(name, "host-emulation (tests only)"); }
long long getMemDeviceCUDA(int) { return 0; }
long long getFreeMemDeviceCUDA(int) { return 0; }
#endif

}  // extern "C"
