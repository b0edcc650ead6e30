"""Parity at the sizes the product runs at (BASELINE configs c1 and c2), through the C ABI on the B200 against the float64 oracle and the
committed c1 golden.  Tolerance: per iteration relL2(gpu, float64 oracle) <= max(4 * eps, 2e-6 * it),
max-abs <= 1e-3 * max(psi)."""
import importlib.util
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
EPS32 = 1.5e-7


def rel_tol(it):
    return max(4 * EPS32, 2e-6 * it)


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_c1_full_ten_iterations_against_oracle_and_golden(product_lib, oracle):
    """config c1 in full: 4 views, 256 x 256 x 128, PSF 25 x 19 x 25, EFFICIENT_BAYESIAN, 10 iterations; every iteration against the
    float64 oracle run beside it, iterations 1 / 2 / 10 and all 40 statistics against tests/golden/c1_case.npz
    (MultiViewDeconvolutionSeq.java:58-180)."""
    import mvrecon_b200 as m
    gen = _load(os.path.join(HERE, "golden", "make_golden_c1.py"), "make_golden_c1")
    gold = np.load(os.path.join(HERE, "golden", "c1_case.npz"))
    T, step = int(gold["quirk_threads"]), int(gold["lattice_step"])
    ds, views, psi0, avg = gen.inputs()
    assert np.allclose([float(im.sum(dtype=np.float64)) for im in ds.images], gold["img_sums"], rtol=1e-12)        # the seeded inputs did not drift
    dv = m.DeconViews([m.DeconView(ds.images[v], ds.weights[v], ds.psfs[v], m.PSFTYPE.EFFICIENT_BAYESIAN) for v in range(4)], norm_quirk_threads=T)
    try:
        for v in range(4):
            assert oracle.rel_l2(dv.views[v].psf.getKernel1(), views[v].kernel1) < 1e-6
            assert oracle.rel_l2(dv.views[v].psf.getKernel2(), views[v].kernel2) < 2e-6
            assert abs(float(dv.views[v].psf.getKernel1().sum(dtype=np.float64)) - gold["k1_sums"][v]) < 1e-6
        dec = m.MultiViewDeconvolutionSeq(dv, 10, m.PsiInitBlurredFused(5.0))
        assert abs(dec.views.lib.dll.mvd_version()) >= 200
        init_psi = dec.getPSI()
        assert oracle.rel_l2(init_psi, psi0) < 1e-6                      # PsiInitBlurredFused on the device == oracle's
        assert np.array_equal(dec.max, gold["max"])
        psi64, k = init_psi, 0                                           # both sides continue from the device's psi0
        for it in range(1, 11):
            stats = dec.runNextIteration()
            for v in range(4):
                psi64, s, mx = oracle.view_update_whole(psi64, views[v], 0.0, dtype=np.float64)
                # sumChange is a SIGNED sum over 8.4 M voxels (DeconvolutionMethods.java:308): its float32 noise floor scales with sum |psi|
                noise = 2e-7 * float(np.abs(psi64).sum(dtype=np.float64))
                assert abs(stats[v].sumChange - s) <= 2e-4 * abs(s) + noise
                assert abs(stats[v].maxChange - mx) <= 2e-3 * abs(mx) + 1e-3
                _, _, s_ref, m_ref = gold["stats"][k]
                assert abs(stats[v].sumChange - s_ref) <= 5e-4 * abs(s_ref) + 2 * noise and abs(stats[v].maxChange - m_ref) <= 2e-3 * abs(m_ref) + 1e-3
                k += 1
            psi = dec.getPSI()
            assert oracle.rel_l2(psi, psi64) <= rel_tol(it), it
            assert np.abs(psi - psi64).max() <= 1e-3 * psi64.max()
            if it in (1, 2, 10):
                ref = gold[f"psi_it{it}"]
                assert oracle.rel_l2(psi[::step, ::step, ::step], ref) <= 2 * rel_tol(it), it        # golden started from the oracle's psi0 (1e-7 apart)
                mom = gold[f"psi_it{it}_moments"]
                p = psi.astype(np.float64)
                assert abs(p.sum() - mom[0]) <= 1e-5 * abs(mom[0]) and abs((p * p).sum() - mom[1]) <= 1e-5 * abs(mom[1])
    finally:
        dv.close()


def test_c2_six_views_tikhonov_two_iterations_against_oracle(product_lib, oracle):
    """config c2 with all six views: 512 x 512 x 256, EFFICIENT_BAYESIAN + Tikhonov (lambda = 0.006), two full iterations on the device.
    Each iteration is checked from the device's own starting state against the float64 oracle on a crop + halo whose core is what
    6 view updates x (k - 1) of contamination leave untouched (the crop spans z completely: true faces there)."""
    import mvrecon_b200 as m
    bench = _load(os.path.join(ROOT, "bench.py"), "bench_for_tests")
    W = bench.WORKLOADS["c2"]
    dims, V = W["dims"], W["views"]
    ds = oracle.make_synthetic(dims, V, seed=20262)
    for v in range(V):
        assert tuple(map(tuple, ds.boxes[v])) == tuple(map(tuple, bench.coverage_box(dims, v)))
    dv = m.DeconViews([m.DeconView(ds.images[v], None, ds.psfs[v], m.PSFTYPE.EFFICIENT_BAYESIAN) for v in range(V)], lambda_=W["lam"])
    try:
        for v in range(V):                                               # weight masks on the device, bit-exact to the oracle's
            mn, mx = ds.boxes[v]
            dv.makeBlendingWeights(v, mn, mx, (0.0,) * 3, (12.0,) * 3)
        dv.normalizeWeights(1.0, False)
        assert np.array_equal(dv.getWeight(3), ds.weights[3])
        init = m.PsiInitBlurredFused(5.0)
        dec = m.MultiViewDeconvolutionSeq(dv, 2, init)
        info = dv.tile_info()
        assert info["num_tiles"] >= 1
        mx = [float(x) for x in init.getMax()]
        region, core = bench.parity_region(W, V, dims[1] // 2, dims[0] // 2)
        rs = tuple(slice(lo, hi) for lo, hi in region)
        cs = tuple(slice(c[0] - r[0], c[1] - r[0]) for c, r in zip(core, region))
        gs = tuple(slice(c[0], c[1]) for c in core)
        imgs = [np.ascontiguousarray(ds.images[v][rs]) for v in range(V)]
        for it in (1, 2):
            before = dec.getPSI()[rs].copy()
            dec.runNextIteration()
            after = dec.getPSI()
            assert np.isfinite(after).all()
            ref = bench.oracle_region_update(W, ds.psfs, region, before, imgs, mx, V)
            assert oracle.rel_l2(after[gs], ref[cs]) <= rel_tol(1), (it, info)
            assert np.abs(after[gs] - ref[cs]).max() <= 1e-3 * np.abs(ref[cs]).max()
    finally:
        dv.close()


def test_psi_init_from_file_on_device(product_lib, oracle, tmp_path):
    """PsiInitFromFile (M/process/deconvolution/init/PsiInitFromFile.java:66-93): psi from a 32-bit TIFF stack, max[] / avg from the precise or the
    approximate average initialiser with setImgToAvg(false); wrong dimensions -> runInitialization returns false."""
    import mvrecon_b200 as m
    ds = oracle.make_synthetic((33, 36, 40), 3, seed=1, psf_size_xyz=(7, 5, 7), psf_sigma_xyz=(1.2, 1.0, 2.0), bead_density=512)
    rng = np.random.default_rng(7)
    start = (50 + 100 * rng.random(ds.dims_zyx)).astype(np.float32)
    path = str(tmp_path / "psi_start.tif")
    product_lib.tiff_write(path, start)
    assert np.array_equal(product_lib.tiff_read(path), start)
    for precise in (True, False):
        dv = m.DeconViews([m.DeconView(ds.images[v], ds.weights[v], ds.psfs[v], m.PSFTYPE.EFFICIENT_BAYESIAN) for v in range(3)])
        try:
            init = m.PsiInitFromFile(path, precise)
            dec = m.MultiViewDeconvolutionSeq(dv, 1, init)
            assert dec.initWasSuccessful()
            assert np.array_equal(dec.getPSI(), start)
            _, mx, avg = (oracle.psi_init_avg_precise if precise else oracle.psi_init_avg_approx)(ds.images, set_img_to_avg=False, psi=start)
            assert np.array_equal(init.getMax(), mx)
            assert (abs(init.getAvg() - avg) <= 1e-9 * abs(avg)) if precise else init.getAvg() == -1.0
            views, _, _ = oracle.make_oracle_views(ds, oracle.EFFICIENT_BAYESIAN)
            for v in range(3):
                views[v].max_intensity = float(mx[v])
            dec.runIterations()
            ref, _ = oracle.run_iterations_seq(start, views, 1, 0.0, dtype=np.float64)
            assert oracle.rel_l2(dec.getPSI(), ref) <= rel_tol(1)
        finally:
            dv.close()
    bad = str(tmp_path / "bad.tif")
    product_lib.tiff_write(bad, start[:-1])
    dv = m.DeconViews([m.DeconView(ds.images[v], ds.weights[v], ds.psfs[v], m.PSFTYPE.EFFICIENT_BAYESIAN) for v in range(3)])
    try:
        assert not m.MultiViewDeconvolutionSeq(dv, 1, m.PsiInitFromFile(bad, True)).initWasSuccessful()
        assert not m.MultiViewDeconvolutionSeq(dv, 1, m.PsiInitFromFile(str(tmp_path / "missing.tif"), True)).initWasSuccessful()
    finally:
        dv.close()


def test_async_upload_with_device_generated_weights(product_lib, oracle, small_dataset):
    """asynchronous view upload + weight == None: the zeroing of the owned weight volume runs on the copy stream behind the image upload and
    must not wipe the masks generated on the compute stream (race in make_blending_weights / normalize_view_weights)."""
    import mvrecon_b200 as m
    ds = small_dataset
    for rep in range(3):
        dv = m.DeconViews([m.DeconView(ds.images[v], None, ds.psfs[v], m.PSFTYPE.EFFICIENT_BAYESIAN) for v in range(3)], async_upload=True)
        try:
            for v in range(3):
                mn, mx = ds.boxes[v]
                dv.makeBlendingWeights(v, mn, mx, (0.0,) * 3, (12.0,) * 3)
            dv.normalizeWeights(1.0, False)
            for v in range(3):
                assert np.array_equal(dv.getWeight(v), ds.weights[v])
        finally:
            dv.close()
