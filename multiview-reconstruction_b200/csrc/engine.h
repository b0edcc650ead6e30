// Host-side engine of the B200 multi-view deconvolution path.
//
// Mirrors the reference's object model for this path (net.preibisch.mvrecon.process.deconvolution):
//   DeconViewPSF  -> kernel1 / kernel2 derivation by PSFTYPE + resident kernel spectra   (DeconViewPSF.java:119-254)
//   DeconView     -> observed image, weight, PSF of one (virtual) view                     (DeconView.java:118-184)
//   DeconViews    -> all views, dimension check, PSF init in list order                    (DeconViews.java:44-81)
//   MultiViewDeconvolutionSeq -> OSEM iteration loop                                       (MultiViewDeconvolutionSeq.java:58-180)
//   Block / BlockGeneratorFixedSizePrecise -> replaced by the TilePlan below: halo'd FFT tiles sized for HBM/L2
//                                             instead of host RAM, same validity rule (BlockGeneratorFixedSizePrecise.java:59-131)
#pragma once
#include <map>
#include <memory>
#include <string>
#include <vector>

#include <functional>

#include "backend.h"

namespace mvd {

enum PsfType : int { OPTIMIZATION_II = 0, OPTIMIZATION_I = 1, EFFICIENT_BAYESIAN = 2, INDEPENDENT = 3 };

// geometry of the real volumes this engine instance works on (all (x,y,z) order)
struct Geometry {
    int gdim[3];     // global fused volume size
    int goff[3];     // global coordinate of local array element 0
    int vol[3];      // local array size
    int own_lo[3];   // responsibility box of this instance (global coords, half open)
    int own_hi[3];
};

struct AxisTile { int org, lo, hi; };
struct AxisTiling { int T = 0; std::vector<AxisTile> tiles; long long cost() const { return (long long)T * (long long)tiles.size(); } };

struct TileGeom { int org[3], lo[3], hi[3]; };

// Reach of a kernel of size k along one axis: output x reads input x - t, t in [-c, k-1-c], c = k/2
struct Reach { int lo, hi; };
inline Reach reach_of(int k) { return Reach{k - 1 - k / 2, k / 2}; }

// device twiddle tables, cached per length
class Tables {
  public:
    explicit Tables(stream_t s) : stream_(s) {}
    ~Tables();
    const cpx* tw(int N);       // per-position stage twiddles of the column passes (LenOps::fill_ctw)
    const cpx* twist(int M);    // exp(-i pi m / (2M))
    const cpx* xtw(int M);      // per-position stage twiddles of the x passes (LenOps::fill_xtw)
  private:
    stream_t stream_;
    std::map<int, cpx*> tw_, twist_, xtw_;
};

// A tiled FFT convolution plan for one geometry and one (or two chained) kernel extents.
class Convolver {
  public:
    // r1: reach of the first convolution, r2: reach of the second (all zero for a single convolution).
    // xmode 0: negacyclic real-packed x axis (halo'd tiles); xmode 1: exact circular convolution at the volume size
    // two_exchanges: the sharded sides of the box only carry max(r1, r2) (scheme B: the host exchanges psi by r1 before the update
    // and the quotient by r2 between the two convolutions, see set_mid_exchange) instead of r1 + r2 (scheme A)
    Convolver(const Geometry& g, const Reach r1[3], const Reach r2[3], int xmode, int max_len, stream_t s, Tables* tables,
              bool two_exchanges = false);
    ~Convolver();
    // called between P5 and P6 of a view update with the tile's x-spectrum of the quotient ([Tz][Ty][px] complex)
    // start: begins the exchange of the quotient's x-spectrum (may return before it completed); join: everything it delivers is in place
    // for the work that follows on the compute stream.  Between the two the pass computes the lines the exchange does not touch.
    void set_mid_exchange(std::function<void(cpx* work, const TileGeom& t)> start, std::function<void()> join = nullptr) {
        mid_exchange_ = std::move(start); mid_join_ = std::move(join);
    }
    int pitch() const { return px_; }

    size_t tile_elems() const { return (size_t)px_ * T_[1] * T_[2]; }            // complex elements per spectrum
    int num_tiles() const { return (int)tiles_.size(); }
    int parts_per_tile() const { return xblocks_; }
    const int* tile_dims() const { return T_; }                                   // {Tx(real), Ty, Tz}
    const std::vector<TileGeom>& tiles() const { return tiles_; }
    double fft_volume_ratio() const;                                              // FFT-box voxels / useful voxels
    int launches_per_update() const {
        const int chunks = chunk_planes_ > 0 ? (T_[2] + chunk_planes_ - 1) / chunk_planes_ : 1;
        return (7 * chunks + 2) * num_tiles();
    }

    // kernel (x fastest, dims kd) -> resident spectrum (scale 1/Nfft folded in); caller owns the buffer
    cpx* build_khat(const float* kernel_host, const int kd[3]);
    void build_khat_into(cpx* khat, const float* kernel_host, const int kd[3]);    // asynchronous on the stream; khat holds tile_elems()
    // dst(own box) = src (*) kernel, src extended by ext
    void conv(const float* src, float* dst, const cpx* khat, int ext, float ext_value);
    // one fused view update (P1..P9) over all tiles; partial stats -> part_sum/part_max [num_tiles*parts_per_tile]
    // skip (may be null): one flag per tile; a flagged tile is not computed -- its part of psi is copied unchanged and it reports no change
    // (a block without content, DeconView.filterBlocksForContent)
    void view_update(const float* psi_in, float* psi_out, const float* img, const float* weight, const cpx* k1hat,
                     const cpx* k2hat, float lambda, float min_value, float max_intensity, double* part_sum, float* part_max,
                     const unsigned char* skip = nullptr, const std::function<void()>* psi_join = nullptr);
    // conv1 -> quotient -> conv2 without the update: the "integral" of one view (P1..P8 + real store), used by the Mul iteration
    void integral(const float* psi_in, const float* img, const cpx* k1hat, const cpx* k2hat, float* integral_out);
    // per-pass CUDA-event timing (bench.py's roofline leg): P1..P9 -> slots 0..8
    void set_profiling(bool on) { prof_on_ = on; }
    void collect_pass_times(double ms[9], long long counts[9], bool reset);
    // what is not a pass: [0] the quotient exchange as the compute stream sees it (start .. next pass), [1] from the end of P9 to the first
    // pass of the next view update (statistics kernels, psi exchange or its start, launch gaps)
    // [2] everything else between passes (joins of a travelling exchange, clearing of rows / planes)
    void collect_aux_times(double ms[3], long long counts[3], bool reset);

  private:
    XArgs base_xargs(const TileGeom& t) const;
    void col(int axis, int mode, const cpx* khat, int z0 = 0, int z1 = -1);
    void xpass(int kind, XArgs a, int z0 = 0, int z1 = -1);
    Geometry g_;
    int xmode_;
    bool two_z_ = false;          // exchange scheme 1 on a z-sharded box: the quotient of the halo planes arrives from the z neighbours
    Reach r2z_ = {0, 0};          // reach of kernel2 along z (planes the quotient exchange fills on either side of the own slab)
    int T_[3];     // real tile extents
    int M_;        // complex x length
    int px_;       // complex pitch
    int xblocks_;
    const LenOps *ox_, *oy_, *oz_;
    std::vector<TileGeom> tiles_;
    stream_t stream_;
    Tables* tables_;
    cpx* work_ = nullptr;
    float* kpad_ = nullptr;
    float* kdev_ = nullptr;     // kernel staging, grows on demand
    size_t kdev_cap_ = 0;
    std::function<void(cpx*, const TileGeom&)> mid_exchange_;
    std::function<void()> mid_join_;
    Reach r2y_ = {0, 0};
    bool shard_lo_[3] = {false, false, false}, shard_hi_[3] = {false, false, false};   // this side of the box has a neighbour (not a volume face)
    // software L2 prefetch distance in CTAs for the x, y and z passes (MVD_PF_X / MVD_PF_Y / MVD_PF_Z override)
    int pf_x_ = 37, pf_y_ = 296, pf_z_ = 0;     // measured (c3, two-stage column plans): the z convolution is 15 % faster without
    int chunk_planes_ = 0;  // planes per L2-resident chunk of the x/y pass chains (0 = whole tile per launch)
    // profiling
    void mark(int pass);
    bool prof_on_ = false;
    std::vector<void*> prof_events_;
    std::vector<int> prof_ids_;
    size_t prof_used_ = 0;
    double prof_ms_[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};      // 0..8: passes P1..P9; 9: quotient exchange; 10: between two view updates
    long long prof_n_[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
};

AxisTiling plan_axis(int gdim, int own_lo, int own_hi, Reach r1, Reach r2, bool is_x, int max_len, bool two_exchanges = false);

// dst = src (*) kernel on the device; host in / host out.  circular: exact circular convolution at the volume size (legacy
// JNA semantics), otherwise U/FFTConvolution semantics (output = size of src, src extended by ext).
void convolve_host(int device, stream_t s, Tables* tables, int max_len, const float* src, const int dims[3], const float* kernel,
                   const int kd[3], int ext, float ext_value, float* dst, bool circular);

struct IterStats { double sum_change; double max_change; };

// in-library halo exchange over NCCL for a (y x z) process grid (comm.cpp)
class NcclComm {                         // one communicator per process and device; reusable across contexts / jobs
  public:
    static void unique_id(char out[128]);
    NcclComm(const char id[128], int world, int rank, int device);
    ~NcclComm();
    int world() const { return world_; }
    int rank() const { return rank_; }
    int device() const { return device_; }
    void* raw() const { return comm_; }
  private:
    void* comm_ = nullptr;
    int world_, rank_, device_;
};
// a [planes][rows][row_floats] float array with this instance's own region inside it, and the widths of the halo to fill below /
// above the own region along y (rows) and z (planes).  Mirrors mvd_halo_box of the C ABI.
struct HaloBox {
    float* base;
    long long row_floats;
    int nrows, nplanes;
    int y0, y1, z0, z1;                  // own region, array indices, half open
    int hy_lo, hy_hi, hz_lo, hz_hi;      // halo rows / planes wanted below (lo) and above (hi) the own region
};
typedef int (*ExchangeFn)(void* user, int which, const HaloBox* box);      // host-provided exchange (mvd_exchange_fn)
typedef int (*ReduceFn)(void* user, double* values, int count, int op);    // host-provided all-reduce, in place (mvd_reduce_fn); op 0 sum, 1 max

class HaloComm {
  public:
    // collective over the communicator.  need_y / need_z: floats this rank pushes per y / z neighbour and exchange at most.
    HaloComm(std::shared_ptr<NcclComm> comm, int py, int pz, stream_t s, size_t need_y, size_t need_z);
    ~HaloComm();
    // enqueued on the stream; the host does not block.  force_nccl: a one-off exchange larger than the landing buffers (PsiInit);
    // every rank must pass the same value
    // s: the stream the exchange is enqueued on (null = the context's compute stream)
    void exchange(const HaloBox& b, bool force_nccl = false, stream_t s = nullptr);
    void all_reduce(double* host_values, int count, int op);      // collective, synchronises the stream; op 0 sum, 1 max
    int transport() const { return peer_ ? 1 : 0; }      // 0: NCCL send/recv, 1: direct stores into the neighbours' memory
  private:
    void reserve(size_t floats, stream_t also = nullptr);
    void exchange_nccl(const HaloBox& b, stream_t s);
    void exchange_peer(const HaloBox& b, stream_t s);
    void setup_peer(size_t need_y, size_t need_z);
    void close_peer();
    std::shared_ptr<NcclComm> comm_;
    int py_, pz_, ry_ = 0, rz_ = 0;
    stream_t stream_;
    size_t stage_floats_ = 0;
    float* stage_[4] = {nullptr, nullptr, nullptr, nullptr};
    double* red_dev_ = nullptr;
    size_t red_cap_ = 0;
    // peer transport: one region per rank = 8 landing buffers ({from lower y, from upper y, from lower z, from upper z} x parity)
    // + 4 arrival counters, mapped into the neighbours with CUDA IPC (or used directly when the neighbour lives in this process)
    bool peer_ = false;
    size_t cap_y_ = 0, cap_z_ = 0;
    float* region_ = nullptr;
    float* nb_region_[4] = {nullptr, nullptr, nullptr, nullptr};     // neighbour regions: lower y, upper y, lower z, upper z
    bool nb_ipc_[4] = {false, false, false, false};
    unsigned seq_ = 0;
    float* landing(float* region, int src, int parity) const;
    unsigned* flags(float* region) const;
};

// ---- point-wise device stages (pointwise.cpp) -------------------------------------------------------------------------
#define MVD_MAX_VIEWS 16
struct ViewPtrs { const float* img[MVD_MAX_VIEWS]; const float* weight[MVD_MAX_VIEWS]; };
struct WeightPtrs { float* w[MVD_MAX_VIEWS]; };
struct MulPtrs { const float* integral[MVD_MAX_VIEWS]; const float* weight[MVD_MAX_VIEWS]; };
// the own (responsibility) region of a local [.][ny][nx] array in array indices, half open
struct OwnBox { int nx, ny; int y0, y1, z0, z1; };
void psi_fused_stats(stream_t s, const ViewPtrs& vp, int V, float* psi, long long n, const OwnBox& ob, double* acc_dev, float* max_dev);
void fill_volume(stream_t s, float* p, long long n, float v);
void slice_stats(stream_t s, const float* img, const OwnBox& ob, long long nyz, double* acc_dev, float* max_dev);
void copy_region(stream_t s, const float* src, int sny, int sy0, int sz0, float* dst, int dny, int dy0, int dz0, int nx, int rows, int planes);
void box_nonzero(stream_t s, const float* w, int nx, int ny, const int lo[3], const int hi[3], int* flag_dev);     // *flag_dev = 1 if any w != 0 in the box
void clear_parts(stream_t s, double* part_sum, float* part_max, int n);                                            // {0, -1} partial statistics
void blend_weights(stream_t s, float* out, const double* lut_dev, const int vol[3], const int goff[3], const int box_min[3], const int box_max[3],
                   const float border[3], const float blending[3], const double* inv_affine = nullptr, const int* bbox_offset = nullptr);
std::vector<double> blend_lut();
// one raw (untransformed) view of a group, resident on the device (pointwise.cpp: FuseGroupKernel)
struct RawViewDev {
    const float* raw; int dims[3];
    double im[12];                       // row-packed inverse of the view -> world model
    int interpolation;                   // 0 nearest neighbour, 1 n-linear
    int fusion_blend; float fusion_border[3], fusion_range[3];
    int decon_blend; float decon_border[3], decon_range[3];
};
void fuse_group(stream_t s, const RawViewDev* views_dev, const RawViewDev* views_host, int count, float* img_out, float* w_out,
                const double* lut_dev, const int vol[3], const int goff[3], const int bbox_min[3], float min_value, float outside_value);
// PSF preparation on the host (kernels are a few thousand voxels; psf_prep.cpp)
void psf_transformed_geometry(const int dims[3], const double affine[12], int new_dims[3], double offset[3]);
std::vector<float> psf_transform_normalized(const float* psf, const int dims[3], const double affine[12], const double inv_affine[12], int new_dims[3]);
std::vector<float> psf_average(const float* const* psfs, const int (*dims)[3], int count, bool use_max, int out_dims[3]);
std::vector<float> psf_make_same_size(const float* psf, const int dims[3], const int new_dims[3]);
// TIFF stacks at the boundary (tiff_io.cpp): PsiInitFromFile / Save3dTIFF
// N5 datasets (csrc/n5_io.cpp): the project's PSF store (psf.n5) and the N5 export of the result; dims (x,y,z)
void n5_dims(const char* dataset_dir, int dims[3]);
std::vector<float> n5_read_f32(const char* dataset_dir, int dims[3]);
void n5_write_f32(const char* dataset_dir, const float* data, const int dims[3], const int block[3], int gzip_level);
void zarr_write_f32(const char* path, const float* data, const int dims[3], const int chunk[3], int gzip_level, const double* voxel_size);
void tiff_dims(const char* path, int dims[3]);
std::vector<float> tiff_read_f32(const char* path, int dims[3]);
void tiff_write_f32(const char* path, const float* data, const int dims[3]);
void normalize_weights(stream_t s, const WeightPtrs& w, int V, long long n, double osem, bool smooth, float max_diff_range, float scaling_range);
void mul_combine(stream_t st, const MulPtrs& p, int V, const float* psi_in, float* psi_out, long long n, long long own0, long long own1, float lambda,
                 float min_value, float max_intensity, double* stats_dev, float* scratch_max_dev);
std::vector<double> gauss3_halfkernel(double sigma);
enum PsiInitType : int { PSI_FUSED_BLURRED = 0, PSI_AVG = 1, PSI_APPROX_AVG = 2 };

class Engine {
  public:
    struct Config {
        int device = 0;
        Geometry geom;
        int num_views = 0;
        int psf_type = EFFICIENT_BAYESIAN;
        float lambda = 0.f;
        float min_value = 1e-4f;
        int max_len = 1152;
        int norm_quirk_threads = 0;   // 0: exact kernel sums; T > 0: reproduce AdjustInput.sumImg's double count for T threads
        int exchange_scheme = 0;      // sharded contexts: 0 = one psi exchange per view update (halo r1 + r2), 1 = psi (r1) + quotient (r2)
    };
    explicit Engine(const Config& c);
    ~Engine();

    int num_views() const { return cfg_.num_views; }
    size_t local_voxels() const { return (size_t)cfg_.geom.vol[0] * cfg_.geom.vol[1] * cfg_.geom.vol[2]; }
    const Config& config() const { return cfg_; }

    void set_view_host(int v, const float* img, const float* weight);
    // asynchronous upload on a copy stream (host buffers should be pinned and must stay valid until the next synchronisation); the
    // first view update of view v waits for its upload only, so the uploads of later views overlap the first iteration
    void set_view_host_async(int v, const float* img, const float* weight);
    void set_view_device(int v, const float* img, const float* weight);
    void set_psf(int v, const float* psf, const int kd[3]);
    void set_kernels(int v, const float* k1, const int k1d[3], const float* k2, const int k2d[3]);
    void init_views();                                     // DeconViews ctor: kernels by PSFTYPE, tile plan, spectra
    void get_kernel_dims(int v, int which, int kd[3]) const;
    void get_kernel(int v, int which, float* out) const;

    void set_psi_host(const float* psi);
    void get_psi_host(float* psi);
    float* psi_device() { join_halo(); return psi_[cur_]; }
    float* psi_next_device() { return psi_[cur_ ^ 1]; }
    void set_max_intensity(const float* mx) { for (int v = 0; v < cfg_.num_views; ++v) views_[v].max_intensity = mx[v]; }
    float max_intensity(int v) const { return views_[v].max_intensity; }

    // PsiInit on the device (PsiInitBlurredFused / PsiInitAvgPrecise / PsiInitAvgApprox): sets psi and the per-view maxima
    // set_img_to_avg = false: only the statistics (PsiInitAvg*.setImgToAvg(false), used by PsiInitFromFile), psi stays as it is
    void psi_init(int type, double sigma, double* avg_out, float* max_out, bool set_img_to_avg = true);
    // weight masks on the device: cosine blending of a view's box, then NormalizingRandomAccess over all views
    void make_blending_weights(int v, const int box_min[3], const int box_max[3], const float border[3], const float blending[3],
                               const double* inv_affine = nullptr, const int* bbox_offset = nullptr);
    void normalize_view_weights(double osem_speedup, bool additional_smooth, float max_diff_range, float scaling_range);
    // ProcessInputImages.fuseGroups for one group: raw host views -> view v's image and (summed) deconvolution weight on the device
    void fuse_group_host(int v, const RawViewDev* views_host /* raw = host pointers */, int count, const int bbox_min[3], float min_value_img,
                         float outside_value);
    void get_weight_host(int v, float* out);
    void get_image_host(int v, float* out);
    double last_fuse_group_ms() const { return last_fuse_ms_; }      // device time of the last fuse_group kernel (CUDA events)
    // MultiViewDeconvolutionMul.runNextIteration: one psi update from all views (geometric mean of the integrals)
    void iteration_mul();
    // attach the NCCL halo exchange: afterwards every view update / Mul iteration is followed by the exchange of the new psi
    void comm_attach(std::shared_ptr<NcclComm> comm, int py, int pz);
    // host-provided exchange instead of NCCL (the stream is synchronised before every call; see mvd_set_exchange_callback)
    void set_exchange_callback(ExchangeFn fn, void* user) { host_exchange_ = fn; host_exchange_user_ = user; install_mid_exchange(); }
    // host-provided all-reduce for the global quantities of a sharded job (per-view maxima, PsiInit average, iteration statistics);
    // with an attached communicator the library uses ncclAllReduce instead
    void set_reduce_callback(ReduceFn fn, void* user) { host_reduce_ = fn; host_reduce_user_ = user; }
    bool is_sharded() const { return sharded(1) || sharded(2); }
    bool can_reduce() const { return host_reduce_ != nullptr || comm_ != nullptr; }
    void all_reduce(double* values, int count, int op);      // no-op on an unsharded context
    void exchange_halos();
    int exchange_transport() const { return host_exchange_ ? 2 : (comm_ ? comm_->transport() : -1); }
    void view_update(int v);                               // asynchronous on the engine stream
    // DeconView.filterBlocksForContent (DeconView.java:204-230): (view, tile) pairs whose weights are all zero inside the tile are not
    // computed.  Returns how many pairs that currently is (evaluates the weights: synchronises; collective on a sharded scheme-1 context,
    // where a tile is only dropped when it is empty on every rank because the neighbours need its quotient).
    int skip_empty_tiles(bool on);
    void fetch_stats(int count, IterStats* out);           // last `count` view updates (synchronises)
    void run_iterations(int n, IterStats* out /* n*V or null */);
    void synchronize() { join_halo(); dev::sync(stream_); }
    stream_t stream() const { return stream_; }

    Convolver* convolver() { return conv_.get(); }
    const Convolver* convolver() const { return conv_.get(); }
    int launches_per_view_update() const;
    // z planes beyond the owned slab whose psi / image values the two chained convolutions actually read
    void halo_needed(int& lo, int& hi) const { lo = halo_lo_; hi = halo_hi_; }
    void halo_needed_y(int& lo, int& hi) const { lo = halo_y_lo_; hi = halo_y_hi_; }

  private:
    struct View {
        const float* img = nullptr;
        const float* weight = nullptr;
        float* img_owned = nullptr;
        float* weight_owned = nullptr;
        std::vector<float> psf, k1, k2;
        int psf_dims[3] = {0, 0, 0}, k1d[3] = {0, 0, 0}, k2d[3] = {0, 0, 0};
        cpx* k1hat = nullptr;
        cpx* k2hat = nullptr;
        float max_intensity = 1.f;
        dev::event_t ready = nullptr;   // upload finished (async uploads)
        bool pending = false;
    };
    stream_t copy_stream_ = nullptr;
    // every use of a view's image / weight on the compute stream first waits for its (asynchronous) upload
    void wait_upload(View& vw) { if (vw.pending) { dev::stream_wait(stream_, vw.ready); vw.pending = false; } }
    std::unique_ptr<Convolver> small_conv_;   // cached plan of the PSF-derivation convolutions
    int small_dims_[3] = {0, 0, 0}, small_kd_[3] = {0, 0, 0};
    float* small_buf_ = nullptr; size_t small_buf_cap_ = 0;      // scratch of conv_same, reused across the derivation
    cpx* small_khat_ = nullptr; size_t small_khat_cap_ = 0;
    void derive_kernels();
    std::vector<float> conv_same(const std::vector<float>& in, const int d[3], const std::vector<float>& k, const int kd[3]);

    Config cfg_;
    stream_t stream_ = nullptr;
    std::unique_ptr<Tables> tables_;
    std::unique_ptr<Convolver> conv_;
    std::vector<View> views_;
    float* psi_[2] = {nullptr, nullptr};
    int cur_ = 0;
    bool inited_ = false;
    int halo_lo_ = 0, halo_hi_ = 0, halo_y_lo_ = 0, halo_y_hi_ = 0;
    // statistics
    double* part_sum_ = nullptr;
    float* part_max_ = nullptr;
    double* stats_dev_ = nullptr;   // ring of {sum,max} pairs
    void ensure_stats_slot();
    std::unique_ptr<HaloComm> comm_;
    double last_fuse_ms_ = 0.0;
    ExchangeFn host_exchange_ = nullptr;
    void* host_exchange_user_ = nullptr;
    ReduceFn host_reduce_ = nullptr;
    void* host_reduce_user_ = nullptr;
    void psi_blur_sharded(const std::vector<float>& k3, int k);
    Reach r1_[3] = {{0, 0}, {0, 0}, {0, 0}}, r2_[3] = {{0, 0}, {0, 0}, {0, 0}};     // kernel reaches (max over views)
    bool sharded(int d) const { return cfg_.geom.own_lo[d] != 0 || cfg_.geom.own_hi[d] != cfg_.geom.gdim[d]; }
    bool has_exchange() const { return comm_ != nullptr || host_exchange_ != nullptr; }
    // ev != null: asynchronous on the exchange stream, `ev` is recorded behind it (the caller joins); else complete on return of the
    // compute stream's point of view
    void do_exchange(int which, const HaloBox& b, bool oversize = false, dev::event_t ev = nullptr);
    // The psi exchange that follows a view update is not waited for: it travels (on its own high-priority stream, or -- host callback --
    // is simply deferred) while the next update transforms the lines that do not read halo data.  join_halo() makes the halos valid for
    // whatever follows on the compute stream; every entry point that touches psi calls it.
    void start_psi_exchange(float* psi);
    void join_halo();
    bool halo_pending_ = false;
    float* pending_psi_ = nullptr;
    bool overlap_ = false;                  // exchanges run on xstream_ (peer transport)
    stream_t xstream_ = nullptr;
    dev::event_t ev_compute_ = nullptr, ev_psi_ = nullptr, ev_mid_ = nullptr;
    void exchange_psi(float* psi);
    HaloBox psi_box(float* psi) const;
    HaloBox spectrum_box(cpx* work, const TileGeom& t) const;

    void install_mid_exchange();
    std::vector<float*> integral_;  // Mul iteration: one integral volume per view
    double* lut_dev_ = nullptr;     // cosine blending LUT
    double* acc_dev_ = nullptr;     // {sum, count} scratch
    float* max_dev_ = nullptr;      // per-view maxima scratch
    int stats_cap_ = 0, stats_count_ = 0;
    bool skip_on_ = false, skip_valid_ = false;
    std::vector<unsigned char> skip_;   // [view][tile]
    int* flag_dev_ = nullptr;
    int refresh_skip();
    void weights_changed() { skip_valid_ = false; }
};

// helpers shared with the C ABI
std::vector<float> mirror_kernel(const std::vector<float>& k, const int kd[3]);        // Mirror.mirror on all axes (Quirk C kept)
double sum_kernel(const std::vector<float>& k);
void norm_to_sum1(std::vector<float>& k, int quirk_threads = 0);

}  // namespace mvd
