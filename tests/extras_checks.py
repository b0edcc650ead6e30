"""Shared checks for PsiInit / weight masks / Mul iteration vs the oracle; run against the host emulator (CPU) and the CUDA library (GPU)."""
import numpy as np
import pytest


def _raw_blend(oracle, ds, v):
    mn, mx = ds.boxes[v]
    return oracle.blending_weight(ds.dims_zyx, mn, mx, (0.0,) * 3, (12.0,) * 3)


def check_weights_on_device_match_oracle(lib, oracle, small_dataset):
    import mvrecon_b200 as m
    ds = small_dataset
    dv = m.DeconViews([m.DeconView(ds.images[v], None, ds.psfs[v], m.PSFTYPE.INDEPENDENT) for v in range(3)], library=lib)
    try:
        for v in range(3):
            mn, mx = ds.boxes[v]
            dv.makeBlendingWeights(v, mn, mx, (0.0,) * 3, (12.0,) * 3)
            assert np.array_equal(dv.getWeight(v), _raw_blend(oracle, ds, v))            # bit exact: same float/double promotion order
        dv.normalizeWeights(1.0, False)
        for v in range(3):
            assert np.array_equal(dv.getWeight(v), ds.weights[v])
    finally:
        dv.close()


def check_weight_normalisation_variants(lib, oracle, small_dataset, smooth, osem):
    import mvrecon_b200 as m
    ds = small_dataset
    raw = [_raw_blend(oracle, ds, v) for v in range(3)]
    ref = oracle.normalize_weights(raw, osem, smooth)
    dv = m.DeconViews([m.DeconView(ds.images[v], raw[v], ds.psfs[v]) for v in range(3)], library=lib)
    try:
        dv.normalizeWeights(osem, smooth)
        for v in range(3):
            assert np.array_equal(dv.getWeight(v), ref[v])
    finally:
        dv.close()


def check_psi_init_variants(lib, oracle, small_dataset):
    import mvrecon_b200 as m
    ds = small_dataset
    mk = lambda: m.DeconViews([m.DeconView(ds.images[v], ds.weights[v], ds.psfs[v]) for v in range(3)], library=lib)
    # FUSED_BLURRED
    ref_psi, ref_max, ref_avg = oracle.psi_init_blurred_fused(ds.images, ds.weights, 5.0)
    dv = mk()
    try:
        init = m.PsiInitBlurredFused(5.0)
        dec = m.MultiViewDeconvolutionSeq(dv, 0, init)
        assert np.array_equal(init.getMax(), ref_max)
        assert abs(init.getAvg() - ref_avg) < 1e-9 * ref_avg
        assert oracle.rel_l2(dec.getPSI(), ref_psi) < 1e-6          # FFT Gaussian vs separable direct sum
    finally:
        dv.close()
    # AVG
    _, ref_max, ref_avg = oracle.psi_init_avg_precise(ds.images)
    dv = mk()
    try:
        init = m.PsiInitAvgPrecise()
        dec = m.MultiViewDeconvolutionSeq(dv, 0, init)
        assert np.array_equal(init.getMax(), ref_max) and abs(init.getAvg() - ref_avg) < 1e-9 * ref_avg
        assert np.all(dec.getPSI() == np.float32(ref_avg))
    finally:
        dv.close()
    # APPROX_AVG (getAvg() == -1 like the reference)
    ref_psi, ref_max, ref_avg = oracle.psi_init_avg_approx(ds.images)
    dv = mk()
    try:
        init = m.PsiInitAvgApprox()
        dec = m.MultiViewDeconvolutionSeq(dv, 0, init)
        assert init.getAvg() == -1.0 and np.array_equal(init.getMax(), ref_max)
        assert np.allclose(dec.getPSI(), ref_psi, rtol=1e-6)
    finally:
        dv.close()
    # no view covers the volume -> error like the reference
    z = np.zeros_like(ds.images[0])
    dv = m.DeconViews([m.DeconView(z, ds.weights[0], ds.psfs[0])], library=lib)
    try:
        with pytest.raises(m.MvdError):
            m.MultiViewDeconvolutionSeq(dv, 0, m.PsiInitBlurredFused())
    finally:
        dv.close()


def check_mul_iteration_matches_oracle(lib, oracle, small_dataset, lam):
    import mvrecon_b200 as m
    ds = small_dataset
    views, psi0, avg = oracle.make_oracle_views(ds, oracle.INDEPENDENT)
    dv = m.DeconViews([m.DeconView(ds.images[v], ds.weights[v], ds.psfs[v], m.PSFTYPE.INDEPENDENT) for v in range(3)], lambda_=lam, library=lib)
    try:
        dec = m.MultiViewDeconvolutionMul(dv, 2, m.PsiInitFromRAI(psi0, [v.max_intensity for v in views]))
        psi = psi0
        for it in range(2):
            st = dec.runNextIteration()[0]
            psi, s, mx = oracle.iteration_mul_whole(psi, views, lam, dtype=np.float64)
            assert oracle.rel_l2(dec.getPSI(), psi) < 4e-6
            assert abs(st.sumChange - s) <= 2e-4 * abs(s) + 0.5 and abs(st.maxChange - mx) <= 2e-3 * abs(mx) + 1e-3
    finally:
        dv.close()


def check_affine_blending_weights(lib, oracle):
    """rotated + scaled view box: device weights == oracle restatement of TransformWeight.transformBlending, bit for bit"""
    import mvrecon_b200 as m
    dims = (30, 36, 44)
    th = np.deg2rad(27.0)
    fwd = np.array([[np.cos(th), 0.0, np.sin(th) * 1.7, 9.3], [0.0, 1.0, 0.0, -2.25], [-np.sin(th), 0.0, np.cos(th) * 1.7, 14.1], [0, 0, 0, 1.0]])
    inv = np.linalg.inv(fwd)[:3].ravel()
    img_min, img_max, off = (0, 0, 0), (39, 35, 19), (-3, 1, 2)
    ref = oracle.blending_weight_affine(dims, off, inv, img_min, img_max, (2.0, 1.0, 0.5), (12.0, 10.0, 6.0))
    assert 0 < np.count_nonzero(ref) < ref.size and ref.max() == 1.0
    img = np.ones(dims, np.float32)
    dv = m.DeconViews([m.DeconView(img, None, np.ones((3, 3, 3), np.float32))], library=lib)
    try:
        dv.makeBlendingWeightsAffine(0, img_min, img_max, inv, off, (2.0, 1.0, 0.5), (12.0, 10.0, 6.0))
        got = dv.getWeight(0)
    finally:
        dv.close()
    assert np.array_equal(got, ref)
    # identity transform reduces to the axis-aligned case
    ident = np.eye(4)[:3].ravel()
    a = oracle.blending_weight_affine(dims, (0, 0, 0), ident, (4, 3, 2), (40, 30, 25), (0, 0, 0), (12, 12, 12))
    b = oracle.blending_weight(dims, (4, 3, 2), (40, 30, 25), (0, 0, 0), (12, 12, 12))
    assert np.array_equal(a, b)
