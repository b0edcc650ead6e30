/*
 * mvdecon.h -- C ABI of the B200-native multi-view deconvolution path.
 *
 * This is the drop-in boundary for ONE hot path of PreibischLab/multiview-reconstruction: the block-wise multi-view
 * Richardson-Lucy / efficient-Bayesian deconvolution loop (net.preibisch.mvrecon.process.deconvolution).
 * Plain C linkage, plain pointers and sizes, no C++/torch types.  Reference citations use
 *     M/ = src/main/java/net/preibisch/mvrecon/      U/ = src/main/java/util/
 *
 * Three levels:
 *   L1  the symbols the UNMODIFIED reference binds through JNA today (CUDAFourierConvolution / CUDAStandardFunctions)
 *   L2  ComputeBlockSeqThread.runIteration on one block (operator boundary)
 *   L3  resident multi-view deconvolution: views, weights, PSF spectra and psi stay in HBM across iterations
 *
 * All volumes are dense float32, x fastest, then y, then z (ImgLib2 ArrayImg order, M/process/cuda/Block.java:299-308).
 * Unless a function says "dims_zyx", dimension triples are in the reference's (x, y, z) order.
 * Every L2/L3 function returns 0 on success; on failure a message is available from mvd_last_error().
 * There is no CPU fallback: every entry point fails if no CUDA device is usable.
 */
#ifndef MVDECON_H
#define MVDECON_H

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define MVD_API __declspec(dllexport)
#else
#define MVD_API __attribute__((visibility("default")))
#endif

/* ------------------------------------------------------------------------------------------------------------------
 * L1 -- legacy JNA boundary (replaces the external FourierConvolutionCUDALib).
 * Java declarations: M/process/cuda/CUDAFourierConvolution.java:26-33, M/process/cuda/CUDAStandardFunctions.java:27-46.
 * Loaded by NativeLibraryTools.loadNativeLibrary (M/process/cuda/NativeLibraryTools.java:84-153), called concurrently from
 * one Java thread per device (M/process/deconvolution/MultiViewDeconvolutionSeq.java:92-150) via
 * ComputeBlockSeqThreadCUDA.convolve1/2 (M/process/deconvolution/iteration/sequential/ComputeBlockSeqThreadCUDA.java:171-208).
 *
 * Semantics: circular convolution at the image size; the kernel is zero padded and its centre sample floor(k/2) is moved to
 * the origin; imDim / kernelDim are {z, y, x} (CUDATools.getCUDACoordinates, M/process/cuda/CUDATools.java:41-49), data x
 * fastest.  The callee never retains the pointers.  No error return (void, as declared by the reference): failures are
 * reported on stderr and leave `im` untouched.  Image dimensions must be supported FFT lengths (every power of two from 32 to
 * 1024 is; the reference GUI enforces powers of two, M/fiji/plugin/fusion/DeconvolutionGUI.java:674-678).
 * ------------------------------------------------------------------------------------------------------------------ */
MVD_API void convolution3DfftCUDAInPlace(float* im, int* imDim, float* kernel, int* kernelDim, int devCUDA);
MVD_API float* convolution3DfftCUDA(float* im, int* imDim, float* kernel, int* kernelDim, int devCUDA); /* malloc'ed result */
MVD_API int getCUDAcomputeCapabilityMinorVersion(int devCUDA);
MVD_API int getCUDAcomputeCapabilityMajorVersion(int devCUDA);
MVD_API int getNumDevicesCUDA(void);                         /* -1: driver/runtime failure (CUDAStandardFunctions.java:39-42) */
MVD_API void getNameDeviceCUDA(int devCUDA, char* name);     /* caller buffer of 256 bytes (CUDATools.java:96-100)           */
MVD_API long long getMemDeviceCUDA(int devCUDA);
MVD_API long long getFreeMemDeviceCUDA(int devCUDA);

/* ------------------------------------------------------------------------------------------------------------------
 * L3 -- resident deconvolution context.
 * Replaces MultiViewDeconvolution / MultiViewDeconvolutionSeq (M/process/deconvolution/MultiViewDeconvolution.java:90-200,
 * MultiViewDeconvolutionSeq.java:58-180), DeconView / DeconViews / DeconViewPSF (DeconView.java:118-184, DeconViews.java:44-81,
 * DeconViewPSF.java:119-254) and the Block machinery (M/process/cuda/Block*.java) for this path.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct mvd_context mvd_context;

/* DeconViewPSF.PSFTYPE ordinals (DeconViewPSF.java:52) */
enum { MVD_PSF_OPTIMIZATION_II = 0, MVD_PSF_OPTIMIZATION_I = 1, MVD_PSF_EFFICIENT_BAYESIAN = 2, MVD_PSF_INDEPENDENT = 3 };
/* out-of-bounds strategies of mvd_convolve */
enum { MVD_EXT_MIRROR = 0, MVD_EXT_ZERO = 1, MVD_EXT_CONST = 2 };

typedef struct mvd_config {
    int device;          /* CUDA device ordinal                                                                         */
    int dims[3];         /* size of the fused / deconvolved volume psi (x,y,z)  -- views.getPSIDimensions()             */
    int num_views;       /* number of (virtual) views                                                                   */
    int psf_type;        /* MVD_PSF_*                                                                                   */
    float lambda;        /* Tikhonov parameter, 0 = off (ComputeBlockSeqThreadCPU.lambda)                               */
    float min_value;     /* MultiViewDeconvolution.minValue = 1e-4f (MultiViewDeconvolution.java:50)                    */
    /* z-slab sharding across processes/GPUs.  Unsharded: shard_lo = 0, shard_hi = dims[2], local_z0 = 0, local_nz = dims[2].
     * Sharded: this context owns planes [shard_lo, shard_hi) and its arrays hold planes [local_z0, local_z0 + local_nz),
     * which must include the halo mvd_halo_planes() reports on interior sides.                                          */
    int shard_lo, shard_hi;
    int local_z0, local_nz;
    int max_fft_len;     /* 0 = default (1152); caps the FFT tile edge                                                  */
    /* AdjustInput.sumImg adds sums[0] and then loops over ALL portion sums including index 0 (AdjustInput.java:115-119), i.e. the
     * first portion (first floor(size/numPortions) samples, numPortions = max(T, size/64^3), T = max(4, #threads),
     * FusionTools.java:1287-1329, Threads.java:40) is counted twice, so the reference's "normalised" kernels do not sum to 1 and
     * the result depends on the thread count.  The default reproduces the reference: 0 = the reference run on THIS host, i.e.
     * T = mvd_reference_threads() = max(4, processors); T > 0 = the reference run with T ImageJ threads (the Java glue passes
     * Threads.numThreads()); < 0 = exact sums (what the author intended; opt-in, differs from the reference by up to ~1 % in the
     * kernel scale and a few percent in psi after 10 iterations on small PSFs).                                                    */
    int norm_quirk_threads;
    /* optional second sharding axis (y), same meaning as the z fields; all zero = y is not sharded.  A 2-d (y x z) process grid keeps the
     * redundant halo volume small: the single-GPU plan already splits y into two FFT tiles, so a 2-way y split costs nothing.            */
    int shard_y_lo, shard_y_hi;
    int local_y0, local_ny;
    /* Halo exchange scheme of a sharded context.  0 (A): one exchange of psi per view update; interior sides carry the reach of BOTH
     * chained convolutions (k1/2 + k2/2) and the box recomputes the neighbour's quotient there.  1 (B): two exchanges -- psi by the
     * reach of kernel1 before the update, and between the two convolutions the quotient (as the rows / planes of its x-spectrum) by
     * the reach of kernel2; interior sides carry max(k1/2, k2/2) only, i.e. less redundant FFT volume.  Scheme 1 needs an attached
     * exchange (mvd_comm_attach or mvd_set_exchange_callback), a box that fits one FFT tile in y and z (several tiles may follow each other
     * along x), and does not support the Mul iteration.                                                                            */
    int exchange_scheme;
} mvd_config;

MVD_API const char* mvd_last_error(void);                     /* thread-local message of the last failing call          */
MVD_API int mvd_version(void);
/* Threads.numThreads() of this host as the library sees it: max(4, processors) (M/Threads.java:40) -- the T that
 * mvd_config.norm_quirk_threads == 0 stands for.                                                                        */
MVD_API int mvd_reference_threads(void);
/* FFT tile lengths compiled into the library (ascending, all 2^a 3^b 5^c); returns the count, fills at most cap entries.              */
MVD_API int mvd_supported_fft_lengths(int* out, int cap);

MVD_API int mvd_create(const mvd_config* cfg, mvd_context** out);
MVD_API int mvd_destroy(mvd_context* ctx);

/* DeconView: observed image (0 where the view has no data, MultiViewDeconvolution.outsideValueImg) and weight volume of
 * view v; local array size dims[0]*dims[1]*local_nz.  _host copies host->device, _device borrows caller-owned device
 * memory that must stay valid until mvd_destroy.  weight == NULL (all three variants): the context allocates the weight
 * volume itself, to be filled on the device by mvd_make_blending_weights* + mvd_normalize_weights ("virtual" weights).   */
MVD_API int mvd_set_view(mvd_context* ctx, int v, const float* img_host, const float* weight_host);
MVD_API int mvd_set_view_device(mvd_context* ctx, int v, const float* img_dev, const float* weight_dev);
/* Asynchronous variant of mvd_set_view: the copies are enqueued on a separate copy stream and the call returns immediately; the host
 * buffers (page-locked for real overlap) must stay valid and unmodified until mvd_synchronize / mvd_get_psi / mvd_run_iterations
 * returned.  The first update of view v waits only for view v's upload, so later uploads overlap the first iteration.                */
MVD_API int mvd_set_view_async(mvd_context* ctx, int v, const float* img_host, const float* weight_host);
/* DeconViewPSF( kernel, psfType ): raw PSF of view v (need not be normalised), size kdims (x,y,z), odd sizes expected.  */
MVD_API int mvd_set_psf(mvd_context* ctx, int v, const float* psf, const int kdims[3]);
/* alternative: hand over kernel1 / kernel2 directly (what runIteration receives, ComputeBlockSeqThread.java:54-61)      */
MVD_API int mvd_set_kernels(mvd_context* ctx, int v, const float* k1, const int k1dims[3], const float* k2, const int k2dims[3]);
/* new DeconViews(...): derives kernel1/kernel2 by PSFTYPE in list order, plans the FFT tiles, builds and keeps all
 * kernel spectra resident in HBM (replaces the lazy DeconViewPSF.getKernel{1,2}FFT, DeconViewPSF.java:79-111).          */
MVD_API int mvd_init_views(mvd_context* ctx);
MVD_API int mvd_get_kernel_dims(mvd_context* ctx, int v, int which /*1|2*/, int kdims[3]);
MVD_API int mvd_get_kernel(mvd_context* ctx, int v, int which /*1|2*/, float* out);

/* psi and the per-view max intensities (what PsiInit hands to the loop, MultiViewDeconvolution.java:115-135)            */
MVD_API int mvd_set_psi(mvd_context* ctx, const float* psi_host);
MVD_API int mvd_get_psi(mvd_context* ctx, float* psi_host);
MVD_API int mvd_set_max_intensities(mvd_context* ctx, const float* max_per_view);

/* PsiInit on the device (M/process/deconvolution/init/PsiInit.java:34 ordinals FUSED_BLURRED=0, AVG=1, APPROX_AVG=2; FROM_FILE / FROM_RAI
 * are mvd_set_psi + mvd_set_max_intensities).  Sets psi and the per-view maxima from the views already handed over
 * (PsiInitBlurredFused.java:63-127, PsiInitAvgPrecise.java:52-112, PsiInitAvgApprox.java:47-99).  sigma: Gaussian of FUSED_BLURRED
 * (reference default 5.0).  avg_out: PsiInit.getAvg() (APPROX_AVG reports -1 like the reference); max_out: num_views values.
 * Fails like the reference when no view covers the volume (FUSED_BLURRED).  On a sharded context the call is collective: avg and the maxima
 * are all-reduced over the attached communicator / reduce callback and the psi halos are exchanged.                                   */
enum { MVD_PSI_FUSED_BLURRED = 0, MVD_PSI_AVG = 1, MVD_PSI_APPROX_AVG = 2 };
MVD_API int mvd_psi_init(mvd_context* ctx, int type, double sigma, double* avg_out, float* max_out);
/* PsiInitFromFile (M/process/deconvolution/init/PsiInitFromFile.java:44-93): psi := the TIFF stack at `path` opened as 32 bit (dimensions
 * must equal the volume, else the call fails like the reference returns false), then avg / max[] from PsiInitAvgPrecise (precise != 0) or
 * PsiInitAvgApprox with setImgToAvg(false).  mvd_tiff_dims / mvd_tiff_read / mvd_tiff_write: the TIFF subset behind it and behind the
 * result export (Save3dTIFF): uncompressed single-channel stacks, 8/16/32-bit integer or 32-bit float in, 32-bit float out with an ImageJ
 * description; dims (x,y,z).  Unsharded contexts only.                                                                              */
MVD_API int mvd_psi_init_from_file(mvd_context* ctx, const char* path, int precise, double* avg_out, float* max_out);
MVD_API int mvd_tiff_dims(const char* path, int dims[3]);
MVD_API int mvd_tiff_read(const char* path, float* out);
MVD_API int mvd_tiff_write(const char* path, const float* data, const int dims[3]);
/* N5 datasets at the same boundary (file-system N5 format restated, github.com/saalfeldlab/n5): dataset_dir = <container>/<dataset>.
 * mvd_n5_dims / mvd_n5_read: what PointSpreadFunction.load does for "psf.n5" (M/fiji/spimdata/pointspreadfunctions/PointSpreadFunction.java:
 * 119-137; N5Utils.open): any 3-d dataset of uint8 / int8 / uint16 / int16 / uint32 / int32 / float32 / float64, raw or gzip, as float32,
 * dims (x,y,z), missing blocks = 0 -- the result feeds mvd_set_psf.  mvd_n5_write: PointSpreadFunction.save (:139-162: block 128^3, gzip
 * level 1) and the N5 export of the deconvolved volume (M/process/export/ExportN5Api.java): float32, gzip_level 0..9 or < 0 = raw.        */
MVD_API int mvd_n5_dims(const char* dataset_dir, int dims[3]);
MVD_API int mvd_n5_read(const char* dataset_dir, float* out);
MVD_API int mvd_n5_write(const char* dataset_dir, const float* data, const int dims[3], const int block_size[3], int gzip_level);
/* OME-Zarr (NGFF 0.4, Zarr v2, one resolution level "0", axes z y x, gzip or raw chunks, "/" separator) export of the result -- the
 * other container ExportN5Api writes; dims / chunk_size (x,y,z), voxel_size (x,y,z) in micrometers or NULL.                          */
MVD_API int mvd_zarr_write(const char* path, const float* data, const int dims[3], const int chunk_size[3], int gzip_level, const double voxel_size[3]);

/* Weight masks on the device.  mvd_make_blending_weights: cosine blending of view v's axis-aligned box [box_min, box_max] (global
 * integer coordinates, inclusive; BlendingRealRandomAccess.computeWeight, M/process/fusion/transformed/weights/BlendingRealRandomAccess.java:95-130)
 * written into the context-owned weight volume of view v (mvd_set_view may be called with weight_host = NULL for such views).
 * mvd_normalize_weights: NormalizingRandomAccess over all views in place (normalization/NormalizingRandomAccess.java:75-109,183-214);
 * reference defaults: osem_speedup 1, additional_smooth 0, max_diff_range 0.1, scaling_range 0.05.  mvd_get_weight: download.        */
MVD_API int mvd_make_blending_weights(mvd_context* ctx, int v, const int box_min[3], const int box_max[3], const float border[3], const float blending[3]);
/* Same for a view whose image interval [img_min, img_max] lives in its own (rotated / scaled) coordinate system:
 * TransformWeight.transformBlending (M/process/fusion/transformed/TransformWeight.java:82-139): every fused voxel (+ bbox_offset, the
 * bounding-box min) is mapped through inv_affine -- the row-packed 3x4 INVERSE of the view -> fused-space AffineTransform3D -- in double,
 * cast to float (TransformedRasteredRandomAccess.applyInverse, .../weights/TransformedRasteredRandomAccess.java:96-114), then weighted.  */
MVD_API int mvd_make_blending_weights_affine(mvd_context* ctx, int v, const int img_min[3], const int img_max[3], const float border[3],
                                             const float blending[3], const double inv_affine[12], const int bbox_offset[3]);
MVD_API int mvd_normalize_weights(mvd_context* ctx, double osem_speedup, int additional_smooth, float max_diff_range, float scaling_range);
MVD_API int mvd_get_weight(mvd_context* ctx, int v, float* weight_host);
MVD_API int mvd_get_image(mvd_context* ctx, int v, float* image_host);      /* download of view v's (possibly generated) image */

/* ---- the step before the loop: view materialisation and PSF preparation -------------------------------------------
 * mvd_fuse_group = ProcessInputImages.fuseGroups for ONE group (M/process/deconvolution/util/ProcessInputImages.java:307-393): every raw
 * view (host memory, zero-min, as the ImgLoader delivers it) is resampled into the context's fused grid on the device --
 * TransformView.transformView (M/process/fusion/transformed/TransformView.java:59-75; TransformedInputRandomAccess.java:60-82): world =
 * voxel + bbox_min, raw position = inv_affine * world in double, n-linear (1) or nearest (0) sample where strictly inside the raw image,
 * max(min_value_img, .) there and outside_value elsewhere (reference: MultiViewDeconvolution.minValueImg = 1, outsideValueImg = 0) -- then
 * image = FusedRandomAccess AVG with the fusion blending weights (FusedRandomAccess.java:66-91), weight = sum of the deconvolution
 * blending weights (CombineWeightsSumRandomAccess.java:40-51).  Blending border / range must already be adjusted
 * (FusionTools.adjustBlending); *_blending = 0 means a constant weight of 1.  Result: view v's image and weight, resident on the device
 * (no fused-size host arrays, no H2D of fused volumes).  Follow with mvd_normalize_weights.                                          */
typedef struct mvd_raw_view {
    const float* raw;
    int dims[3];                 /* (x,y,z) */
    double inv_affine[12];       /* row-packed inverse of the view -> world AffineTransform3D (after any downsampling adjustment) */
    int interpolation;           /* 0 nearest neighbour, 1 linear */
    int fusion_blending; float fusion_border[3], fusion_range[3];
    int decon_blending;  float decon_border[3], decon_range[3];
} mvd_raw_view;
MVD_API int mvd_fuse_group(mvd_context* ctx, int v, const mvd_raw_view* views, int count, const int bbox_min[3], float min_value_img,
                           float outside_value);
MVD_API int mvd_last_fuse_group_ms(mvd_context* ctx, double* ms);     /* device time of the last mvd_fuse_group kernel (CUDA events) */
/* PSFPreparation.loadGroupTransformPSFs (M/process/deconvolution/util/PSFPreparation.java:41-89), host side (PSFs are tiny):
 * mvd_psf_transformed_dims / mvd_psf_transform = PSFExtraction.getTransformedNormalizedPSF (M/process/psf/PSFExtraction.java:182-193,
 * 367-451, 453-473): min-max normalisation, then n-linear resampling of the zero-extended PSF so that the centre stays the centre and all
 * sizes are odd; affine / inv_affine = the view's model and its inverse, row-packed.  mvd_psf_average = PSFCombination.computeAverageImage
 * (M/process/psf/PSFCombination.java:74-135; dims = count x 3 ints; out may be NULL to query out_dims).  mvd_psf_make_same_size =
 * PSFCombination.makeSameSize (:182-212).  The results feed mvd_set_psf.                                                              */
MVD_API int mvd_psf_transformed_dims(const int dims[3], const double affine[12], int new_dims[3]);
MVD_API int mvd_psf_transform(const float* psf, const int dims[3], const double affine[12], const double inv_affine[12], float* out);
MVD_API int mvd_psf_average(const float* const* psfs, const int* dims, int count, int use_max, int out_dims[3], float* out);
MVD_API int mvd_psf_make_same_size(const float* psf, const int dims[3], const int new_dims[3], float* out);

/* MultiViewDeconvolutionMul.runNextIteration (M/process/deconvolution/MultiViewDeconvolutionMul.java:116-245,
 * iteration/mul/ComputeBlockMulThreadCPU.java:87-188, mul/DeconvolutionMethods.java:370-419): ONE psi update per iteration from all
 * views (geometric mean of the per-view integrals, summed weights capped at 1, max := mean of the per-view maxima).                 */
MVD_API int mvd_run_iteration_mul(mvd_context* ctx, double stats[2]);

/* One view update of MultiViewDeconvolutionSeq.runNextIteration (:69-176): psi <- update(psi, view v).
 * stats (may be NULL) receives {sumChange, maxChange} over the owned voxels (IterationStatistics, ComputeBlockThread.java:64-68). */
MVD_API int mvd_run_view_update(mvd_context* ctx, int v, double stats[2]);
/* DeconView.filterBlocksForContent (M/process/deconvolution/DeconView.java:204-274): on != 0 -> a (view, tile) pair whose weight volume
 * is zero everywhere inside the tile is not computed any more (psi keeps its values there, the pair reports sumChange 0 / maxChange -1
 * like a removed block).  The weights are examined now (the call synchronises) and again after every later change of a weight volume;
 * skipped_out (may be NULL) receives the number of pairs currently dropped.  On a sharded context with exchange scheme 1 the call is
 * collective and a tile is dropped only when it is empty on every rank (the neighbours read its quotient).  Off by default.            */
MVD_API int mvd_skip_empty_tiles(mvd_context* ctx, int on, int* skipped_out);
/* MultiViewDeconvolution.runIterations for n iterations; stats (may be NULL) receives n*num_views pairs.                */
MVD_API int mvd_run_iterations(mvd_context* ctx, int n, double* stats);
/* Asynchronous variant + explicit synchronisation (lets a multi-GPU host overlap its halo exchange).                    */
MVD_API int mvd_enqueue_view_update(mvd_context* ctx, int v);
MVD_API int mvd_synchronize(mvd_context* ctx);
MVD_API int mvd_fetch_stats(mvd_context* ctx, int count, double* stats);

/* Introspection for benchmarks / multi-GPU hosts */
/* The tile planner on its own (csrc/engine.cpp plan_axis): how an axis of the fused volume of size gdim, of which this context owns
 * [own_lo, own_hi), is cut into FFT tiles when the two chained kernels reach (r1_lo, r1_hi) and (r2_lo, r2_hi) samples below / above a
 * voxel.  is_x: the x axis (real-packed, tile lengths 2 * FFT length); two_exchanges: exchange scheme 1 on a sharded axis.  Returns the tile
 * length and up to cap tiles as triples {origin, valid_lo, valid_hi} (global coordinates).  Hosts use it to size shards (sharding.py).   */
MVD_API int mvd_plan_axis(int gdim, int own_lo, int own_hi, int r1_lo, int r1_hi, int r2_lo, int r2_hi, int is_x, int max_fft_len,
                          int two_exchanges, int* tile_len, int* tiles, int cap, int* num_tiles);
MVD_API int mvd_tile_info(mvd_context* ctx, int tile_dims[3], int* num_tiles, double* fft_volume_ratio, int* launches_per_view_update);
MVD_API int mvd_halo_planes(mvd_context* ctx, int* lo, int* hi);            /* z planes needed beyond the owned slab       */
MVD_API int mvd_halo_rows(mvd_context* ctx, int* lo, int* hi);              /* y rows needed beyond the owned box (y sharding) */
MVD_API int mvd_psi_device_ptr(mvd_context* ctx, void** current);           /* device address of the current psi buffer    */
MVD_API int mvd_stream_handle(mvd_context* ctx, void** cuda_stream);

/* Multi-GPU: one process (or context) per GPU on a py x pz (y x z) grid of boxes, rank = ry * pz + rz (mvd_config.shard_* describe the
 * box).  mvd_comm_unique_id: rank 0 creates a 128-byte NCCL id that the host distributes with its own plumbing.  mvd_comm_create
 * (collective) builds a communicator for `device` once per process -- it is reusable across contexts and jobs -- and
 * mvd_comm_attach (after mvd_init_views) attaches it to a context.  From then on every view update / Mul iteration is followed by the
 * halo exchange of the new psi (y rows first, then z planes including the fresh y halos), enqueued on the context's stream -- so
 * mvd_run_iterations works unchanged across GPUs.  mvd_exchange_halos triggers one exchange explicitly (e.g. after mvd_psi_init).
 * NCCL is loaded with dlopen on first use; single-GPU use needs no NCCL.                                                              */
typedef struct mvd_comm mvd_comm;
MVD_API int mvd_comm_unique_id(char id_out[128]);
MVD_API int mvd_comm_create(const char id[128], int world, int rank, int device, mvd_comm** out);
MVD_API int mvd_comm_destroy(mvd_comm* comm);
MVD_API int mvd_comm_attach(mvd_context* ctx, mvd_comm* comm, int py, int pz);
MVD_API int mvd_exchange_halos(mvd_context* ctx);
/* How the halos travel: -1 no exchange attached, 0 NCCL send/recv, 1 direct stores into the neighbours' HBM over NVLink (CUDA IPC /
 * peer access; chosen automatically when every rank can map its neighbours, MVD_EXCHANGE=nccl forces 0), 2 host callback.          */
MVD_API int mvd_exchange_transport(mvd_context* ctx, int* transport);
/* Host-provided exchange instead of NCCL (a JVM copying between its contexts with cudaMemcpyPeer, MPI, the CPU tests).  The box is a
 * [nplanes][nrows][row_floats] float array in the context's memory space with the own region rows [y0,y1) x planes [z0,z1); the
 * callback must fill hy_lo rows below / hy_hi rows above the own rows over the own planes from the y neighbours, THEN hz_lo / hz_hi
 * whole planes (all rows, fresh y halos included) from the z neighbours -- and serve the neighbours symmetrically: the neighbour below
 * wants this box's first h*_hi own rows / planes, the one above its last h*_lo.  which: 0 = psi, 1 = x-spectrum of the quotient
 * (scheme 1).  The context's stream is synchronised before the call; the data must be in place when the callback returns 0.          */
typedef struct mvd_halo_box {
    float* base;
    long long row_floats;
    int nrows, nplanes;
    int y0, y1, z0, z1;
    int hy_lo, hy_hi, hz_lo, hz_hi;
} mvd_halo_box;
typedef int (*mvd_exchange_fn)(void* user, int which, const mvd_halo_box* box);
MVD_API int mvd_set_exchange_callback(mvd_context* ctx, mvd_exchange_fn fn, void* user);
/* Global quantities of a sharded job -- the per-view maxima and the average of PsiInit (MultiViewDeconvolution.java:115-135), the
 * IterationStatistics of a view update over all blocks (MultiViewDeconvolutionSeq.java:165-176) -- are all-reduced over the attached
 * communicator (ncclAllReduce), or through this host callback when the host brings its own plumbing: reduce values[0..count) in place
 * over all ranks, op 0 = sum, 1 = max; return 0 on success.  Calls that return such quantities (mvd_psi_init, mvd_psi_init_from_file,
 * mvd_run_iterations with stats, mvd_run_view_update with stats, mvd_fetch_stats) are collective on a sharded context.               */
typedef int (*mvd_reduce_fn)(void* user, double* values, int count, int op);
MVD_API int mvd_set_reduce_callback(mvd_context* ctx, mvd_reduce_fn fn, void* user);

/* Per-pass device timing (CUDA events on the context's stream around every pass launch): slots 0..8 = passes P1..P9 of a
 * view update.  ms[] are accumulated milliseconds, counts[] the number of launches; reset != 0 clears them.   */
MVD_API int mvd_set_profiling(mvd_context* ctx, int on);
MVD_API int mvd_get_pass_times(mvd_context* ctx, double ms[9], long long counts[9], int reset);
/* The rest of a view update as the compute stream sees it: [0] the quotient exchange of scheme 1 (from its start to the next pass),
 * [1] from the end of P9 to the first pass of the next view update (statistics kernels, psi exchange, launch gaps), [2] everything else
 * between passes (joining a travelling exchange, clearing rows / planes nobody delivers).                                             */
MVD_API int mvd_get_aux_times(mvd_context* ctx, double ms[3], long long counts[3], int reset);

/* Generic FFT convolution of a host volume with a host kernel on the device (used by the PSF derivation; exported for
 * tests and callers that need U/FFTConvolution.convolve semantics, U/FFTConvolution.java:490-603): out has the size of img,
 * kernel centre floor(k/2), img extended by `ext`.                                                                      */
MVD_API int mvd_convolve(int device, const float* img, const int dims[3], const float* kernel, const int kdims[3], int ext,
                         float ext_value, float* out);

/* ------------------------------------------------------------------------------------------------------------------
 * L2 -- ComputeBlockSeqThread.runIteration on ONE halo'd block held in host memory
 * (M/process/deconvolution/iteration/sequential/ComputeBlockSeqThread.java:54-61; CPU variant ComputeBlockSeqThreadCPU.java:79-169).
 * psi_block is read and overwritten (it is the worker's psiBlockTmp), img_block / weight_block are the zero-extended block
 * cut-outs, block_dims (x,y,z).  conv1 extends the block by mirroring, conv2 by the constant 1 exactly like the CPU thread.
 * stats receives {sumChange, maxChange} over the whole block (as the reference does).
 * ------------------------------------------------------------------------------------------------------------------ */
MVD_API int mvd_block_iteration(int device, float* psi_block, const float* img_block, const float* weight_block,
                                const int block_dims[3], const float* kernel1, const int k1dims[3], const float* kernel2,
                                const int k2dims[3], float lambda, float min_value, float max_intensity, double stats[2]);

#ifdef __cplusplus
}
#endif
#endif /* MVDECON_H */
