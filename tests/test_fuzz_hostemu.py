"""Randomised shapes through the host-emulated kernel bodies vs the oracle: odd / even / unit kernel extents, volumes smaller than the
PSF, forced multi-tile plans, every PSF type, with and without Tikhonov.  One OSEM iteration per example, compared per view update."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st


COUNTS = {"ran": 0, "rejected": 0, "multitile": 0}


@settings(max_examples=150, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(st.data())
def test_random_geometries_one_iteration(hostemu_lib, oracle, data):
    import mvrecon_b200 as m
    dims = tuple(data.draw(st.integers(5, 36)) for _ in range(3))                    # z, y, x
    kd = tuple(data.draw(st.integers(1, 7)) for _ in range(3))                       # PSF extents z, y, x (even sizes included)
    V = data.draw(st.integers(1, 3))
    ptype = data.draw(st.sampled_from([0, 1, 2, 3]))
    lam = data.draw(st.sampled_from([0.0, 0.02]))
    max_len = data.draw(st.sampled_from([0, 32, 48]))
    seed = data.draw(st.integers(0, 10 ** 6))
    rng = np.random.default_rng(seed)
    psfs = [(0.05 + rng.random(kd)).astype(np.float32) for _ in range(V)]
    imgs = [(1.0 + 50.0 * rng.random(dims)).astype(np.float32) for _ in range(V)]
    for im in imgs:                                                                  # holes: no image data -> quotient 1
        im[rng.random(dims) < 0.1] = 0.0
    ws = [rng.random(dims).astype(np.float32) / V for _ in range(V)]
    psi0 = (5.0 + 20.0 * rng.random(dims)).astype(np.float32)
    mx = [float(im.max()) for im in imgs]
    k1, k2 = oracle.derive_kernels(psfs, ptype)
    views = [oracle.OracleView(imgs[v], ws[v], k1[v], k2[v], mx[v]) for v in range(V)]
    try:
        dv = m.DeconViews([m.DeconView(imgs[v], ws[v], psfs[v], m.PSFTYPE(ptype)) for v in range(V)], lambda_=lam, max_fft_len=max_len,
                          library=hostemu_lib)
    except m.MvdError as e:
        assert "no supported FFT length fits" in str(e) or "larger than the FFT tile" in str(e), str(e)
        COUNTS["rejected"] += 1
        return
    COUNTS["ran"] += 1
    COUNTS["multitile"] += dv.tile_info()["num_tiles"] > 1
    try:
        for v in range(V):
            assert oracle.rel_l2(dv.getViews()[v].psf.getKernel2(), k2[v]) < 2e-5
        dec = m.MultiViewDeconvolutionSeq(dv, 0, m.PsiInitFromRAI(psi0, mx))
        ref = psi0
        for v in range(V):
            st_ = dec.lib.dll.mvd_enqueue_view_update(dv._ctx, v)
            assert st_ == 0
            dv.synchronize()
            ref, _, _ = oracle.view_update_whole(ref, views[v], lam, dtype=np.float64)
            got = dec.getPSI()
            assert np.isfinite(got).all()
            assert oracle.rel_l2(got, ref) < 5e-6, (dims, kd, V, ptype, lam, max_len, seed, v)
    finally:
        dv.close()


def test_zz_fuzz_coverage():
    """runs after the fuzz test of this module: most examples must really have been compared, some of them multi-tile"""
    if COUNTS["ran"] + COUNTS["rejected"] == 0:
        pytest.skip("fuzz test not run")
    print(COUNTS)
    assert COUNTS["ran"] >= 3 * max(COUNTS["rejected"], 1) and COUNTS["multitile"] >= 3, COUNTS
