"""GPU parity tests proper: the shipped CUDA library, called through its C ABI, against the CPU oracle on the same seeded inputs,
against the committed golden vectors, and -- at BASELINE sizes -- through size-independent properties.
Tolerance: relL2(gpu, float64 oracle) <= max(4*eps_it, 2e-6*it), max-abs <= 1e-3 * max(psi)."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "small_case.npz")
EPS32 = 1.5e-7


def rel_tol(it):            # it = 1-based iteration
    return max(4 * EPS32, 2e-6 * it)


def _views(m, ds, ptype, n=None):
    n = n or len(ds.psfs)
    return [m.DeconView(ds.images[v], ds.weights[v], ds.psfs[v], m.PSFTYPE(ptype)) for v in range(n)]


def test_library_loaded_is_the_in_tree_cuda_build(product_lib):
    import mvrecon_b200 as m
    assert os.path.samefile(product_lib.path, m.LIBRARY_PATH)
    assert product_lib.getNumDevicesCUDA() >= 1
    assert product_lib.dll.getCUDAcomputeCapabilityMajorVersion(0) == 10


@pytest.mark.parametrize("shape,ks", [((9, 10, 11), (3, 5, 3)), ((40, 50, 70), (7, 5, 9)), ((130, 100, 90), (25, 19, 25)),
                                      ((33, 65, 129), (4, 6, 2)), ((5, 6, 7), (7, 5, 7)), ((1, 40, 40), (1, 5, 5))])
@pytest.mark.parametrize("ext", ["mirror", "zero", "const"])
def test_convolve_matches_oracle(product_lib, oracle, shape, ks, ext):
    rng = np.random.default_rng(sum(shape))
    img = rng.random(shape).astype(np.float32)
    k = rng.random(ks).astype(np.float32)
    got = product_lib.convolve(img, k, ext, ext_value=1.0)
    ref = oracle.fft_convolve(img, k, ext, const=1.0, dtype=np.float64)
    assert oracle.rel_l2(got, ref) < 5e-7
    assert np.abs(got - ref).max() < 1e-5 * np.abs(ref).max()


@pytest.mark.parametrize("shape", [(32, 32, 32), (64, 32, 128), (256, 256, 256), (40, 36, 60)])
def test_legacy_convolution3DfftCUDAInPlace(product_lib, oracle, shape):
    """L1: circular convolution, kernel centre at the origin, dims {z,y,x}  (ComputeBlockSeqThreadCUDA.java:171-208)."""
    rng = np.random.default_rng(1)
    img = rng.random(shape).astype(np.float32)
    k = rng.random((5, 7, 9)).astype(np.float32)
    ref = oracle.circular_convolve(img, k, dtype=np.float64)
    got = img.copy()
    product_lib.convolution3DfftCUDAInPlace(got, k, 0)
    assert oracle.rel_l2(got, ref) < 5e-7


def test_legacy_out_of_place_and_device_queries(product_lib, oracle):
    rng = np.random.default_rng(2)
    img = rng.random((32, 40, 48)).astype(np.float32)
    k = rng.random((3, 3, 5)).astype(np.float32)
    imd, kd = (C.c_int * 3)(*img.shape), (C.c_int * 3)(*k.shape)
    ptr = product_lib.dll.convolution3DfftCUDA(img.ctypes.data_as(C.POINTER(C.c_float)), imd, k.ctypes.data_as(C.POINTER(C.c_float)), kd, 0)
    assert ptr
    got = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), shape=img.shape).copy()
    C.CDLL(None).free(C.c_void_p(ptr))
    assert oracle.rel_l2(got, oracle.circular_convolve(img, k, dtype=np.float64)) < 5e-7
    assert "B200" in product_lib.getNameDeviceCUDA(0) or len(product_lib.getNameDeviceCUDA(0)) > 0
    assert product_lib.dll.getFreeMemDeviceCUDA(0) > 0


@pytest.mark.parametrize("ptype", [0, 1, 2, 3])
@pytest.mark.parametrize("lam", [0.0, 0.006])
def test_loop_parity_per_iteration(product_lib, oracle, small_dataset, ptype, lam):
    import mvrecon_b200 as m
    ds = small_dataset
    views, psi0, avg = oracle.make_oracle_views(ds, ptype)
    dv = m.DeconViews(_views(m, ds, ptype), lambda_=lam)
    try:
        for v in range(3):
            assert oracle.rel_l2(dv.views[v].psf.getKernel1(), views[v].kernel1) < 1e-6
            assert oracle.rel_l2(dv.views[v].psf.getKernel2(), views[v].kernel2) < 2e-6
        dec = m.MultiViewDeconvolutionSeq(dv, 10, m.PsiInitFromRAI(psi0, [v.max_intensity for v in views]))
        psi64 = psi0
        for it in range(1, 11):
            stats = dec.runNextIteration()
            for v in range(3):
                psi64, s, mx = oracle.view_update_whole(psi64, views[v], lam, dtype=np.float64)
                assert abs(stats[v].sumChange - s) <= 2e-4 * abs(s) + 1.0
                assert abs(stats[v].maxChange - mx) <= 2e-3 * abs(mx) + 1e-3
            psi = dec.getPSI()
            assert oracle.rel_l2(psi, psi64) <= rel_tol(it), it
            assert np.abs(psi - psi64).max() <= 1e-3 * psi64.max()
    finally:
        dv.close()


def test_golden_vectors(product_lib, oracle):
    import mvrecon_b200 as m
    gold = np.load(GOLD)
    views = [m.DeconView(gold[f"img{v}"], gold[f"weight{v}"], gold[f"psf{v}"], m.PSFTYPE.EFFICIENT_BAYESIAN) for v in range(3)]
    dv = m.DeconViews(views, lambda_=float(gold["lambda"]), norm_quirk_threads=int(gold["quirk_threads"]))
    try:
        for ptype in (2,):
            for v in range(3):
                assert oracle.rel_l2(dv.views[v].psf.getKernel1(), gold[f"k1_t{ptype}_v{v}"]) < 1e-6
                assert oracle.rel_l2(dv.views[v].psf.getKernel2(), gold[f"k2_t{ptype}_v{v}"]) < 2e-6
        dec = m.MultiViewDeconvolutionSeq(dv, 5, m.PsiInitFromRAI(gold["psi0"], gold["max"]))
        k = 0
        for it in range(5):
            for v in range(3):
                st = (C.c_double * 2)()
                assert dv.lib.dll.mvd_run_view_update(dv._ctx, v, st) == 0
                if it < 2:
                    assert oracle.rel_l2(dec.getPSI(), gold[f"psi_it{it}_v{v}"]) <= rel_tol(it + 1)
                _, _, s_ref, m_ref = gold["stats"][k]
                assert abs(st[0] - s_ref) <= 2e-4 * abs(s_ref) + 0.5 and abs(st[1] - m_ref) <= 2e-3 * abs(m_ref) + 1e-3
                k += 1
        assert oracle.rel_l2(dec.getPSI(), gold["psi_it4"]) <= rel_tol(5)
    finally:
        dv.close()


@pytest.mark.parametrize("ptype", [0, 1, 2, 3])
def test_golden_kernels_all_psf_types(product_lib, oracle, ptype):
    import mvrecon_b200 as m
    gold = np.load(GOLD)
    views = [m.DeconView(gold[f"img{v}"], gold[f"weight{v}"], gold[f"psf{v}"], m.PSFTYPE(ptype)) for v in range(3)]
    dv = m.DeconViews(views, norm_quirk_threads=int(gold["quirk_threads"]))
    try:
        for v in range(3):
            assert oracle.rel_l2(dv.views[v].psf.getKernel1(), gold[f"k1_t{ptype}_v{v}"]) < 1e-6
            assert oracle.rel_l2(dv.views[v].psf.getKernel2(), gold[f"k2_t{ptype}_v{v}"]) < 2e-6
    finally:
        dv.close()


@pytest.mark.parametrize("lib_t,ora_t", [(0, "reference"), (8, 8), (-1, None)])
def test_norm_quirk_switch(product_lib, oracle, small_dataset, lib_t, ora_t):
    """AdjustInput.sumImg's double count (AdjustInput.java:115-119): the library DEFAULT (0) is the reference's behaviour on this host
    (T = Threads.numThreads()), T > 0 a reference run with T threads, -1 the exact sums; each against the oracle's same switch."""
    import mvrecon_b200 as m
    assert product_lib.dll.mvd_reference_threads() == oracle.num_threads()
    views, psi0, avg = oracle.make_oracle_views(small_dataset, oracle.EFFICIENT_BAYESIAN, quirk_threads=ora_t)
    dv = m.DeconViews(_views(m, small_dataset, 2), lambda_=0.006, norm_quirk_threads=lib_t)
    try:
        for v in range(3):
            assert oracle.rel_l2(dv.views[v].psf.getKernel1(), views[v].kernel1) < 1e-6
        dec = m.MultiViewDeconvolutionSeq(dv, 3, m.PsiInitFromRAI(psi0, [v.max_intensity for v in views]))
        dec.runIterations()
        ref, _ = oracle.run_iterations_seq(psi0, views, 3, 0.006, dtype=np.float64)
        assert oracle.rel_l2(dec.getPSI(), ref) <= rel_tol(3)
    finally:
        dv.close()


@pytest.mark.parametrize("max_len", [40, 48, 64, 0])
def test_tiling_independence(product_lib, oracle, max_len):
    """halo'd multi-tile plans reproduce the whole-volume result (MultiViewDeconvolutionSeq's blocks == whole volume)."""
    import mvrecon_b200 as m
    ds = oracle.make_synthetic((70, 90, 100), 2, seed=2, psf_size_xyz=(9, 7, 9), psf_sigma_xyz=(1.5, 1.2, 2.5), bead_density=2048)
    views, psi0, avg = oracle.make_oracle_views(ds, oracle.EFFICIENT_BAYESIAN)
    dv = m.DeconViews(_views(m, ds, 2), max_fft_len=max_len)
    try:
        dec = m.MultiViewDeconvolutionSeq(dv, 2, m.PsiInitFromRAI(psi0, [v.max_intensity for v in views]))
        dec.runIterations()
        psi = dec.getPSI()
        info = dv.tile_info()
    finally:
        dv.close()
    ref, _ = oracle.run_iterations_seq(psi0, views, 2, 0.0, dtype=np.float64)
    assert oracle.rel_l2(psi, ref) <= rel_tol(2), info


def test_block_operator_and_blocked_driver(product_lib, oracle, small_dataset):
    """L2: ComputeBlockSeqThread.runIteration on halo'd blocks, driven like MultiViewDeconvolutionSeq.runNextIteration."""
    import mvrecon_b200 as m
    ds = small_dataset
    views, psi0, avg = oracle.make_oracle_views(ds, oracle.INDEPENDENT)
    mx = [v.max_intensity for v in views]
    fac = m.ComputeBlockSeqThreadB200Factory(1e-4, 0.006, (32, 32, 32), devices=(0,))
    assert fac.numParallelBlocks() == 1
    psi = psi0.copy()
    mv = [m.DeconView(ds.images[v], ds.weights[v], ds.psfs[v]) for v in range(3)]
    stats = m.runNextIterationBlocked(psi, mv, [(v.kernel1, v.kernel2) for v in views], mx, fac)
    ref, st = oracle.run_iterations_seq(psi0, views, 1, 0.006, dtype=np.float64)
    refb, stb = oracle.run_iterations_seq(psi0, views, 1, 0.006, dtype=np.float64, block_size_xyz=(32, 32, 32))
    assert oracle.rel_l2(psi, ref) <= rel_tol(1)
    for v in range(3):      # block statistics include the halo, exactly like the reference (ComputeBlockSeqThreadCPU.java:129-157)
        assert abs(stats[v].sumChange - stb[v][2]) <= 2e-4 * abs(stb[v][2]) + 1.0


def test_sharded_contexts_equal_single_context(product_lib, oracle):
    import mvrecon_b200 as m
    ds = oracle.make_synthetic((64, 36, 40), 2, seed=3, psf_size_xyz=(5, 5, 7), psf_sigma_xyz=(1.0, 1.0, 1.6), bead_density=1024)
    views, psi0, avg = oracle.make_oracle_views(ds, oracle.EFFICIENT_BAYESIAN)
    mx = [v.max_intensity for v in views]
    nz, H = 64, 6
    shards = []
    for lo, hi in [(0, 20), (20, 47), (47, 64)]:
        z0, z1 = max(0, lo - H), min(nz, hi + H)
        loc = [m.DeconView(ds.images[v][z0:z1], ds.weights[v][z0:z1], ds.psfs[v], m.PSFTYPE.EFFICIENT_BAYESIAN) for v in range(2)]
        d = m.DeconViews(loc, shard=(lo, hi, z0, z1 - z0), global_dims_zyx=(nz, 36, 40))
        shards.append((d, m.MultiViewDeconvolutionSeq(d, 0, m.PsiInitFromRAI(psi0[z0:z1], mx)), lo, hi, z0, z1))
    full = psi0.copy()
    psi64 = psi0
    for it in range(2):
        for v in range(2):
            new = np.empty_like(full)
            for d, dec, lo, hi, z0, z1 in shards:
                d.enqueue_view_update(v)
                d.synchronize()
                new[lo:hi] = dec.getPSI()[lo - z0:hi - z0]
            full = new
            psi64, _, _ = oracle.view_update_whole(psi64, views[v], 0.0, dtype=np.float64)
            assert oracle.rel_l2(full, psi64) <= rel_tol(it + 1)
            for d, dec, lo, hi, z0, z1 in shards:
                assert d.lib.dll.mvd_set_psi(d._ctx, np.ascontiguousarray(full[z0:z1]).ctypes.data_as(m._F)) == 0
    for d, *_ in shards:
        d.close()


def test_edge_cases(product_lib, oracle):
    import mvrecon_b200 as m
    rng = np.random.default_rng(5)
    # single view falls back to the flipped kernel; weight 0 leaves psi bit-exact; img == 0 -> quotient 1
    shape = (20, 24, 28)
    psf = oracle.synth_psf(0, 1, (5, 5, 5), (1.0, 1.0, 1.0))
    img = (50 + 10 * rng.random(shape)).astype(np.float32)
    img[:, :, :5] = 0
    w = np.ones(shape, np.float32)
    w[:6] = 0
    psi0 = (40 + rng.random(shape)).astype(np.float32)
    dv = m.DeconViews([m.DeconView(img, w, psf, m.PSFTYPE.EFFICIENT_BAYESIAN)])
    try:
        k1, k2 = dv.views[0].psf.getKernel1(), dv.views[0].psf.getKernel2()
        assert np.array_equal(k2, k1[::-1, ::-1, ::-1])
        dec = m.MultiViewDeconvolutionSeq(dv, 1, m.PsiInitFromRAI(psi0, [60.0]))
        dec.runIterations()
        psi = dec.getPSI()
    finally:
        dv.close()
    assert np.array_equal(psi[:6], psi0[:6])
    ref, _, _ = oracle.view_update_whole(psi0, oracle.OracleView(img, w, k1, k2, 60.0), 0.0, dtype=np.float64)
    assert oracle.rel_l2(psi, ref) <= rel_tol(1)
    # value <= 0 / NaN -> minValue: psi with zeros and an image that makes the blur vanish
    psiz = np.zeros(shape, np.float32)
    dv = m.DeconViews([m.DeconView(img, np.ones(shape, np.float32), psf)])
    try:
        dec = m.MultiViewDeconvolutionSeq(dv, 1, m.PsiInitFromRAI(psiz, [60.0]))
        dec.runIterations()
        out = dec.getPSI()
    finally:
        dv.close()
    assert np.all(out == np.float32(1e-4))
    # errors surface as exceptions with a message, never as silent fallbacks
    with pytest.raises(m.MvdError):
        m.DeconViews([m.DeconView(img, w, psf), m.DeconView(img[:-1], w[:-1], psf)])
    with pytest.raises(m.MvdError):
        m.DeconViews([m.DeconView(img, w, np.ones((5, 5, 5000), np.float32))])        # PSF larger than any FFT tile


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE-size properties (config c2: 512 x 512 x 256): size-independent invariants instead of a CPU recomputation
# ---------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def c2_volume():
    rng = np.random.default_rng(11)
    return (100 + 50 * rng.random((256, 512, 512), dtype=np.float32)).astype(np.float32)


def test_fullsize_convolution_linearity_and_mass(product_lib, oracle, c2_volume):
    psf = oracle.synth_psf(1, 6)
    a = c2_volume
    b = np.roll(a, 7, axis=2)
    ca, cb = product_lib.convolve(a, psf, "mirror"), product_lib.convolve(b, psf, "mirror")
    cab = product_lib.convolve((2 * a + 3 * b).astype(np.float32), psf, "mirror")
    assert oracle.rel_l2(cab, 2 * ca.astype(np.float64) + 3 * cb) < 1e-6
    const = np.full(a.shape, 7.0, np.float32)
    assert np.abs(product_lib.convolve(const, psf, "mirror") - 7.0).max() < 5e-5           # sum K = 1
    # zero extension conserves mass: sum(out) over an enlarged support == sum(in) * sum(K); check the interior identity instead
    delta = np.zeros((3, 3, 3), np.float32)
    delta[1, 1, 1] = 1
    assert np.array_equal(product_lib.convolve(a, delta, "mirror"), a) or oracle.rel_l2(product_lib.convolve(a, delta, "mirror"), a) < 3e-7


def test_fullsize_update_properties(product_lib, oracle, c2_volume):
    """6-view 512x512x256 EB + Tikhonov (config c2): determinism, weight-0 invariance, agreement with the oracle on a sub-block."""
    import mvrecon_b200 as m
    V = 3                                             # three of the six views keep host memory modest
    psfs = [oracle.synth_psf(v, 6) for v in range(V)]
    img = c2_volume
    w = np.full(img.shape, 1.0 / V, np.float32)
    w[:, :, 300:] = 0
    views = [m.DeconView(np.roll(img, 3 * v, axis=1), w, psfs[v], m.PSFTYPE.EFFICIENT_BAYESIAN) for v in range(V)]
    outs = []
    for rep in range(2):
        dv = m.DeconViews(views, lambda_=0.006)
        try:
            dec = m.MultiViewDeconvolutionSeq(dv, 1, m.PsiInitFromRAI(img, [200.0] * V))
            dec.runIterations()
            outs.append(dec.getPSI())
            k = [(dv.views[v].psf.getKernel1(), dv.views[v].psf.getKernel2()) for v in range(V)]
        finally:
            dv.close()
    assert np.array_equal(outs[0], outs[1])                                     # bitwise deterministic
    assert np.array_equal(outs[0][:, :, 300:], img[:, :, 300:])                 # weight 0 => untouched
    assert np.isfinite(outs[0]).all()
    # oracle on a sub-block far from the volume faces: the halo makes the crop exact (block independence)
    z, y, x, s, h = 100, 200, 100, 48, 30
    sl = (slice(z - h, z + s + h), slice(y - h, y + s + h), slice(x - h, x + s + h))
    psi = img[sl].astype(np.float32)
    for v in range(V):
        ov = oracle.OracleView(np.roll(img, 3 * v, axis=1)[sl], w[sl], k[v][0], k[v][1], 200.0)
        nxt, _, _ = oracle.view_update_whole(psi, ov, 0.006, dtype=np.float64)
        # only the centre of the crop is valid after each view update; shrink the trusted region by the kernel reach
        psi = nxt
    core = (slice(h, h + s),) * 3
    got = outs[0][z:z + s, y:y + s, x:x + s]
    # after 3 view updates the crop is contaminated up to 3*(k-1) = 72 voxels from its faces in z/x; compare one update only
    dv = m.DeconViews(views[:1], lambda_=0.006)
    try:
        dec = m.MultiViewDeconvolutionSeq(dv, 1, m.PsiInitFromRAI(img, [200.0]))
        dec.runIterations()
        one = dec.getPSI()
        k1, k2 = dv.views[0].psf.getKernel1(), dv.views[0].psf.getKernel2()
    finally:
        dv.close()
    ov = oracle.OracleView(img[sl], w[sl], k1, k2, 200.0)
    ref, _, _ = oracle.view_update_whole(img[sl], ov, 0.006, dtype=np.float64)
    assert oracle.rel_l2(one[z:z + s, y:y + s, x:x + s], ref[core]) <= rel_tol(1)
    assert got.shape == (s, s, s)


# ---------------------------------------------------------------------------------------------------------------------
# scaled-down versions of the BASELINE configs c4 / c5 (the full sizes need 8 GPUs): many views, large anisotropic PSFs
# ---------------------------------------------------------------------------------------------------------------------
def test_c5_like_large_anisotropic_psf_optimization_ii(product_lib, oracle):
    """7 views, OPTIMIZATION_II, anisotropic 15x15x31 PSFs tilted by 4 degree steps, 3 iterations (config c5 scaled by 1/2 in the PSF)."""
    import mvrecon_b200 as m
    ds = oracle.make_synthetic((96, 72, 80), 7, seed=20265, psf_size_xyz=(15, 15, 31), psf_sigma_xyz=(1.3, 1.3, 4.5), bead_density=4096,
                               tilt_step_deg=4.0)
    views, psi0, avg = oracle.make_oracle_views(ds, oracle.OPTIMIZATION_II)
    dv = m.DeconViews(_views(m, ds, oracle.OPTIMIZATION_II))
    try:
        info = dv.tile_info()
        dec = m.MultiViewDeconvolutionSeq(dv, 3, m.PsiInitFromRAI(psi0, [v.max_intensity for v in views]))
        psi64 = psi0
        for it in range(1, 4):
            dec.runNextIteration()
            for v in range(7):
                psi64, _, _ = oracle.view_update_whole(psi64, views[v], 0.0, dtype=np.float64)
            assert oracle.rel_l2(dec.getPSI(), psi64) <= rel_tol(it), (it, info)
    finally:
        dv.close()


def test_c4_like_independent_eight_views_device_weights(product_lib, oracle):
    """8 views, INDEPENDENT (classic multi-view RL, kernel2 = flipped kernel1), weight masks generated and normalised on the device."""
    import mvrecon_b200 as m
    ds = oracle.make_synthetic((48, 64, 72), 8, seed=20264, psf_size_xyz=(9, 7, 9), psf_sigma_xyz=(1.4, 1.2, 2.2), bead_density=2048)
    views, psi0, avg = oracle.make_oracle_views(ds, oracle.INDEPENDENT)
    dv = m.DeconViews([m.DeconView(ds.images[v], None, ds.psfs[v], m.PSFTYPE.INDEPENDENT) for v in range(8)])
    try:
        for v in range(8):
            mn, mx = ds.boxes[v]
            dv.makeBlendingWeights(v, mn, mx, (0.0,) * 3, (12.0,) * 3)
        dv.normalizeWeights(1.0, False)
        for v in range(8):
            assert np.array_equal(dv.getWeight(v), ds.weights[v])
        dec = m.MultiViewDeconvolutionSeq(dv, 2, m.PsiInitBlurredFused(5.0))
        assert oracle.rel_l2(dec.getPSI(), psi0) < 1e-6
        dec.runIterations()
        ref, _ = oracle.run_iterations_seq(psi0, views, 2, 0.0, dtype=np.float64)
        assert oracle.rel_l2(dec.getPSI(), ref) <= 3 * rel_tol(2)       # psi0 differs at 1e-7 (FFT Gaussian vs separable sum)
    finally:
        dv.close()


def test_async_upload_equals_synchronous(product_lib, oracle, small_dataset):
    import mvrecon_b200 as m
    views, psi0, avg = oracle.make_oracle_views(small_dataset, oracle.EFFICIENT_BAYESIAN)
    outs = []
    for asyn in (False, True):
        dv = m.DeconViews(_views(m, small_dataset, 2), lambda_=0.006, async_upload=asyn)
        try:
            dec = m.MultiViewDeconvolutionSeq(dv, 2, m.PsiInitFromRAI(psi0, [v.max_intensity for v in views]))
            dec.runIterations()
            buf = np.empty_like(psi0)
            outs.append(dec.getPSI(out=buf).copy())
        finally:
            dv.close()
    assert np.array_equal(outs[0], outs[1])


def test_two_devices_in_one_process(product_lib, oracle):
    """the reference's multi-GPU model: one process, one worker per device (ComputeBlockSeqThreadCUDAFactory.java:57-64)"""
    if product_lib.getNumDevicesCUDA() < 2:
        pytest.skip("needs two devices")
    import threading
    rng = np.random.default_rng(9)
    img = rng.random((64, 48, 80)).astype(np.float32)
    k = rng.random((9, 7, 5)).astype(np.float32)
    ref = oracle.fft_convolve(img, k, "mirror", dtype=np.float64)
    out = {}

    def work(dev):
        out[dev] = product_lib.convolve(img, k, "mirror", device=dev)

    ts = [threading.Thread(target=work, args=(d,)) for d in (0, 1)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    for d in (0, 1):
        assert oracle.rel_l2(out[d], ref) < 5e-7


@pytest.mark.gpu
@pytest.mark.parametrize("axis,scheme", [("z", 0), ("y", 0), ("z", 1), ("y", 1)])
def test_in_library_halo_exchange_two_gpus(product_lib, oracle, axis, scheme):
    """two contexts on two devices joined by the library's own NCCL exchange (mvd_comm_create / mvd_comm_attach): mvd_run_iterations
    on each shard, without any host-side exchange, must equal the whole-volume update"""
    if product_lib.getNumDevicesCUDA() < 2:
        pytest.skip("needs two devices")
    import threading
    import torch  # noqa: F401  (loads torch's bundled NCCL before the library dlopens one: a process can hold only one libnccl.so.2)
    import mvrecon_b200 as m
    ds = oracle.make_synthetic((48, 40, 36), 2, seed=5, psf_size_xyz=(5, 7, 5), psf_sigma_xyz=(1.0, 1.4, 1.2), bead_density=1024)
    views, psi0, avg = oracle.make_oracle_views(ds, oracle.EFFICIENT_BAYESIAN)
    mx = [v.max_intensity for v in views]
    nz, ny = 48, 40
    n = nz if axis == "z" else ny
    cut = n // 2 + 3
    H = (3 if scheme == 1 else 6) if axis == "y" else (2 if scheme == 1 else 4)        # PSF 5 x 7 x 5 (x, y, z)
    uid = product_lib.comm_unique_id()
    res, err = {}, []

    def work(r):
        try:
            lo, hi = (0, cut) if r == 0 else (cut, n)
            a0, a1 = max(0, lo - H), min(n, hi + H)
            sl = (slice(a0, a1),) if axis == "z" else (slice(None), slice(a0, a1))
            loc = [m.DeconView(np.ascontiguousarray(ds.images[v][sl]), np.ascontiguousarray(ds.weights[v][sl]), ds.psfs[v],
                               m.PSFTYPE.EFFICIENT_BAYESIAN) for v in range(2)]
            kw = {"shard": (lo, hi, a0, a1 - a0)} if axis == "z" else {"shard_y": (lo, hi, a0, a1 - a0)}
            d = m.DeconViews(loc, global_dims_zyx=(nz, ny, 36), device=r, exchange_scheme=scheme, **kw)
            assert (d.halo_planes() if axis == "z" else d.halo_rows()) == ((0, H) if r == 0 else (H, 0))
            comm = product_lib.comm_create(uid, 2, r, r)
            d.comm_attach(comm, *((1, 2) if axis == "z" else (2, 1)))
            dec = m.MultiViewDeconvolutionSeq(d, 2, m.PsiInitFromRAI(np.ascontiguousarray(psi0[sl]), mx))
            dec.runIterations()
            own = (slice(lo - a0, hi - a0),) if axis == "z" else (slice(None), slice(lo - a0, hi - a0))
            res[r] = (lo, hi, dec.getPSI()[own].copy())
            d.close()
            comm.close()
        except Exception as e:  # noqa: BLE001
            err.append(e)

    ts = [threading.Thread(target=work, args=(r,), daemon=True) for r in (0, 1)]
    [t.start() for t in ts]
    [t.join(timeout=120) for t in ts]
    assert not err, err
    assert len(res) == 2
    psi64 = psi0
    for it in range(2):
        for v in range(2):
            psi64, _, _ = oracle.view_update_whole(psi64, views[v], 0.0, dtype=np.float64)
    full = np.empty_like(psi0)
    for r in (0, 1):
        lo, hi, part = res[r]
        if axis == "z":
            full[lo:hi] = part
        else:
            full[:, lo:hi] = part
    assert oracle.rel_l2(full, psi64) <= rel_tol(2)


def _two_process_worker(rank, world, port, out_dir, axis, scheme, transport):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["MVD_EXCHANGE"] = transport
    import torch
    import torch.distributed as dist
    import mvdecon_oracle as o
    import mvrecon_b200 as m
    from mvrecon_b200 import sharding
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)       # plumbing for the unique id only
    lib = m.lib()
    dims = (48, 40, 36)
    ds = o.make_synthetic(dims, 2, seed=5, psf_size_xyz=(5, 7, 5), psf_sigma_xyz=(1.0, 1.4, 1.2), bead_density=1024)
    views, psi0, avg = o.make_oracle_views(ds, o.EFFICIENT_BAYESIAN)
    n = dims[0] if axis == "z" else dims[1]
    H = (3 if scheme == 1 else 6) if axis == "y" else (2 if scheme == 1 else 4)
    lo, hi = sharding.slab_range(n, world, rank)
    a0, a1 = sharding.extended_range(lo, hi, n, H)
    sl = (slice(a0, a1),) if axis == "z" else (slice(None), slice(a0, a1))
    loc = [m.DeconView(np.ascontiguousarray(ds.images[v][sl]), np.ascontiguousarray(ds.weights[v][sl]), ds.psfs[v],
                       m.PSFTYPE.EFFICIENT_BAYESIAN) for v in range(2)]
    kw = {"shard": (lo, hi, a0, a1 - a0)} if axis == "z" else {"shard_y": (lo, hi, a0, a1 - a0)}
    d = m.DeconViews(loc, global_dims_zyx=dims, device=rank, exchange_scheme=scheme, **kw)
    ids = [lib.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    comm = lib.comm_create(ids[0], world, rank, rank)
    d.comm_attach(comm, *((1, world) if axis == "z" else (world, 1)))
    assert d.exchange_transport() == ("peer-stores" if transport == "peer" else "nccl")
    dec = m.MultiViewDeconvolutionSeq(d, 3, m.PsiInitFromRAI(np.ascontiguousarray(psi0[sl]), [v.max_intensity for v in views]))
    dec.runIterations()
    own = (slice(lo - a0, hi - a0),) if axis == "z" else (slice(None), slice(lo - a0, hi - a0))
    np.save(os.path.join(out_dir, f"part{rank}.npy"), dec.getPSI()[own])
    d.close()
    comm.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("axis,scheme,transport", [("y", 1, "peer"), ("z", 1, "peer"), ("z", 0, "peer"), ("y", 1, "nccl")])
def test_two_processes_two_gpus(product_lib, oracle, tmp_path, axis, scheme, transport):
    """one process per GPU (the bench.py / production layout): halos stored straight into the neighbour's HBM through CUDA IPC
    mappings, or sent with NCCL; three iterations must equal the whole-volume oracle"""
    if product_lib.getNumDevicesCUDA() < 2:
        pytest.skip("needs two devices")
    import torch.multiprocessing as mp
    port = 32500 + (os.getpid() % 2000) + 11 * scheme + (5 if axis == "y" else 0) + (23 if transport == "nccl" else 0)
    mp.start_processes(_two_process_worker, args=(2, port, str(tmp_path), axis, scheme, transport), nprocs=2, join=True, start_method="spawn")
    ds = oracle.make_synthetic((48, 40, 36), 2, seed=5, psf_size_xyz=(5, 7, 5), psf_sigma_xyz=(1.0, 1.4, 1.2), bead_density=1024)
    views, psi0, avg = oracle.make_oracle_views(ds, oracle.EFFICIENT_BAYESIAN)
    ref, _ = oracle.run_iterations_seq(psi0, views, 3, 0.0, dtype=np.float64)
    got = np.concatenate([np.load(tmp_path / f"part{r}.npy") for r in range(2)], axis=0 if axis == "z" else 1)
    assert oracle.rel_l2(got, ref) <= rel_tol(3)
