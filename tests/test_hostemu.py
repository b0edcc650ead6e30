"""CPU-side validation of the CUDA kernel bodies through the host emulator (same source, threads run sequentially).
This checks index math, tiling, kernel placement and the pass schedule against the oracle without a GPU.
The emulator is test infrastructure only -- the product library has no CPU path."""
import numpy as np
import pytest


def _views(m, ds, ptype, n=None):
    n = n or len(ds.psfs)
    return [m.DeconView(ds.images[v], ds.weights[v], ds.psfs[v], m.PSFTYPE(ptype)) for v in range(n)]


@pytest.mark.parametrize("shape,ks", [((9, 10, 11), (3, 5, 3)), ((20, 33, 47), (5, 4, 7)), ((40, 41, 30), (9, 7, 5))])
@pytest.mark.parametrize("ext", ["mirror", "zero", "const"])
def test_convolve_matches_oracle(hostemu_lib, oracle, shape, ks, ext):
    rng = np.random.default_rng(7)
    img = rng.random(shape).astype(np.float32)
    k = rng.random(ks).astype(np.float32)
    got = hostemu_lib.convolve(img, k, ext, ext_value=1.0)
    ref = oracle.fft_convolve(img, k, ext, const=1.0, dtype=np.float64)
    assert oracle.rel_l2(got, ref) < 5e-7


def test_legacy_circular(hostemu_lib, oracle):
    rng = np.random.default_rng(8)
    img = rng.random((32, 36, 40)).astype(np.float32)
    k = rng.random((5, 3, 7)).astype(np.float32)
    ref = oracle.circular_convolve(img, k, dtype=np.float64)
    got = img.copy()
    hostemu_lib.convolution3DfftCUDAInPlace(got, k, 0)
    assert oracle.rel_l2(got, ref) < 5e-7


@pytest.mark.parametrize("ptype", [0, 1, 2, 3])
def test_loop_parity_all_psf_types(hostemu_lib, oracle, small_dataset, ptype):
    import mvrecon_b200 as m
    ds = small_dataset
    views, psi0, avg = oracle.make_oracle_views(ds, ptype)
    dv = m.DeconViews(_views(m, ds, ptype), lambda_=0.006, library=hostemu_lib)
    try:
        for v in range(3):
            assert oracle.rel_l2(dv.views[v].psf.getKernel1(), views[v].kernel1) < 1e-6
            assert oracle.rel_l2(dv.views[v].psf.getKernel2(), views[v].kernel2) < 2e-6
        dec = m.MultiViewDeconvolutionSeq(dv, 2, m.PsiInitFromRAI(psi0, [v.max_intensity for v in views]))
        dec.runIterations()
        psi = dec.getPSI()
    finally:
        dv.close()
    p64, st = oracle.run_iterations_seq(psi0, views, 2, 0.006, dtype=np.float64)
    assert oracle.rel_l2(psi, p64) < 4e-6
    assert np.abs(psi - p64).max() < 1e-3 * p64.max()
    for i, (it, v, s, mx) in enumerate(st):
        got = dec.stats[it][v]
        assert abs(got.sumChange - s) <= 1e-4 * max(abs(s), 1.0) + 0.5
        assert abs(got.maxChange - mx) <= 1e-3 * max(abs(mx), 1.0)


def test_norm_quirk_switch_matches_oracle(hostemu_lib, oracle, small_dataset):
    """mvd_config.norm_quirk_threads reproduces AdjustInput.sumImg's double count exactly like the oracle's switch."""
    import mvrecon_b200 as m
    ds = small_dataset
    assert hostemu_lib.dll.mvd_reference_threads() == oracle.num_threads()
    # (library setting, oracle setting): the DEFAULT of both is the reference's own behaviour on this host; -1 / None = exact sums
    for lib_t, ora_t in ((0, oracle.REFERENCE_THREADS), (8, 8), (5, 5), (-1, None)):
        views, psi0, avg = oracle.make_oracle_views(ds, oracle.EFFICIENT_BAYESIAN, quirk_threads=ora_t)
        dv = m.DeconViews(_views(m, ds, 2), norm_quirk_threads=lib_t, library=hostemu_lib)
        try:
            for v in range(3):
                assert abs(float(dv.views[v].psf.getKernel1().sum(dtype=np.float64)) - float(views[v].kernel1.sum(dtype=np.float64))) < 1e-6
                assert oracle.rel_l2(dv.views[v].psf.getKernel1(), views[v].kernel1) < 1e-6
                assert oracle.rel_l2(dv.views[v].psf.getKernel2(), views[v].kernel2) < 2e-6
            if lib_t == -1:
                assert abs(float(dv.views[0].psf.getKernel1().sum(dtype=np.float64)) - 1.0) < 1e-6
            else:
                assert abs(float(dv.views[0].psf.getKernel1().sum(dtype=np.float64)) - 1.0) > 1e-4        # the reference's kernels do NOT sum to 1
        finally:
            dv.close()


def test_multitile_equals_single_tile(hostemu_lib, oracle):
    """several halo'd tiles per axis (max_fft_len forces tiling) give the whole-volume result."""
    import mvrecon_b200 as m
    ds = oracle.make_synthetic((50, 60, 70), 2, seed=2, psf_size_xyz=(7, 5, 7), psf_sigma_xyz=(1.3, 1.1, 2.0), bead_density=1024)
    views, psi0, avg = oracle.make_oracle_views(ds, oracle.EFFICIENT_BAYESIAN)
    dv = m.DeconViews(_views(m, ds, 2), lambda_=0.0, max_fft_len=40, library=hostemu_lib)
    try:
        info = dv.tile_info()
        assert info["num_tiles"] >= 4
        dec = m.MultiViewDeconvolutionSeq(dv, 1, m.PsiInitFromRAI(psi0, [v.max_intensity for v in views]))
        dec.runIterations()
        psi = dec.getPSI()
    finally:
        dv.close()
    p64, _ = oracle.run_iterations_seq(psi0, views, 1, 0.0, dtype=np.float64)
    assert oracle.rel_l2(psi, p64) < 4e-6


def test_block_iteration_matches_reference_block(hostemu_lib, oracle, small_dataset):
    """L2 operator: one halo'd block with mirror / constant-1 extension == ComputeBlockSeqThreadCPU.runIteration."""
    ds = small_dataset
    views, psi0, avg = oracle.make_oracle_views(ds, oracle.INDEPENDENT)
    v = views[0]
    ref, s, mx = oracle.view_update_whole(psi0, v, 0.006, dtype=np.float64)
    blk = psi0.copy()
    gs, gm = hostemu_lib.block_iteration(blk, v.image, v.weight, v.kernel1, v.kernel2, 0.006, 1e-4, v.max_intensity)
    assert oracle.rel_l2(blk, ref) < 2e-6
    assert abs(gs - s) <= 1e-4 * abs(s) + 0.5


def test_sharded_matches_unsharded(hostemu_lib, oracle):
    """z-slab sharding with halo planes supplied by the host == single context (scheme A)."""
    import mvrecon_b200 as m
    ds = oracle.make_synthetic((64, 36, 40), 2, seed=3, psf_size_xyz=(5, 5, 7), psf_sigma_xyz=(1.0, 1.0, 1.6), bead_density=1024)
    views, psi0, avg = oracle.make_oracle_views(ds, oracle.EFFICIENT_BAYESIAN)
    mx = [v.max_intensity for v in views]
    nz = 64
    # reference: one context
    dv = m.DeconViews(_views(m, ds, 2), library=hostemu_lib)
    dec = m.MultiViewDeconvolutionSeq(dv, 1, m.PsiInitFromRAI(psi0, mx))
    ref_after = []
    for v in range(2):
        st = dec.lib.dll.mvd_run_view_update(dv._ctx, v, None)
        assert st == 0
        ref_after.append(dec.getPSI())
    dv.close()
    # two shards; halo = kz - 1 on the interior side
    H = 6
    cuts = [(0, 32), (32, 64)]
    shards = []
    for lo, hi in cuts:
        z0, z1 = max(0, lo - H), min(nz, hi + H)
        loc = [m.DeconView(ds.images[v][z0:z1], ds.weights[v][z0:z1], ds.psfs[v], m.PSFTYPE.EFFICIENT_BAYESIAN) for v in range(2)]
        d = m.DeconViews(loc, shard=(lo, hi, z0, z1 - z0), global_dims_zyx=(nz, 36, 40), library=hostemu_lib)
        hl, hh = d.halo_planes()
        assert hl <= lo - z0 and hh <= z1 - hi
        shards.append((d, m.MultiViewDeconvolutionSeq(d, 1, m.PsiInitFromRAI(psi0[z0:z1], mx)), lo, hi, z0, z1))
    full = psi0.copy()
    for v in range(2):
        new = np.empty_like(full)
        for d, dec, lo, hi, z0, z1 in shards:
            assert dec.lib.dll.mvd_run_view_update(d._ctx, v, None) == 0
            new[lo:hi] = dec.getPSI()[lo - z0:hi - z0]
        full = new
        assert oracle.rel_l2(full, ref_after[v]) < 1e-6
        # halo exchange through the host: hand every shard the updated extended slab
        for d, dec, lo, hi, z0, z1 in shards:
            assert dec.lib.dll.mvd_set_psi(d._ctx, np.ascontiguousarray(full[z0:z1]).ctypes.data_as(m._F)) == 0
    for d, *_ in shards:
        d.close()


def test_2d_sharding_y_and_z_matches_unsharded(hostemu_lib, oracle):
    """(y x z) = 2 x 2 process grid emulated in one process: every shard owns a box, halos are refreshed from the assembled volume."""
    import mvrecon_b200 as m
    dims = (40, 44, 36)
    ds = oracle.make_synthetic(dims, 2, seed=4, psf_size_xyz=(5, 5, 5), psf_sigma_xyz=(1.0, 1.0, 1.2), bead_density=1024)
    views, psi0, avg = oracle.make_oracle_views(ds, oracle.EFFICIENT_BAYESIAN)
    mx = [v.max_intensity for v in views]
    nz, ny, nx = dims
    H = 4
    shards = []
    for (ylo, yhi) in [(0, 22), (22, 44)]:
        for (zlo, zhi) in [(0, 20), (20, 40)]:
            y0, y1, z0, z1 = max(0, ylo - H), min(ny, yhi + H), max(0, zlo - H), min(nz, zhi + H)
            loc = [m.DeconView(np.ascontiguousarray(ds.images[v][z0:z1, y0:y1]), np.ascontiguousarray(ds.weights[v][z0:z1, y0:y1]),
                               ds.psfs[v], m.PSFTYPE.EFFICIENT_BAYESIAN) for v in range(2)]
            d = m.DeconViews(loc, shard=(zlo, zhi, z0, z1 - z0), shard_y=(ylo, yhi, y0, y1 - y0), global_dims_zyx=dims, library=hostemu_lib)
            assert d.halo_rows() == ((0 if ylo == 0 else H), (0 if yhi == ny else H))
            dec = m.MultiViewDeconvolutionSeq(d, 0, m.PsiInitFromRAI(np.ascontiguousarray(psi0[z0:z1, y0:y1]), mx))
            shards.append((d, dec, (ylo, yhi, y0, y1), (zlo, zhi, z0, z1)))
    full = psi0.copy()
    psi64 = psi0
    for it in range(2):
        for v in range(2):
            new = np.empty_like(full)
            for d, dec, (ylo, yhi, y0, y1), (zlo, zhi, z0, z1) in shards:
                d.enqueue_view_update(v)
                d.synchronize()
                new[zlo:zhi, ylo:yhi] = dec.getPSI()[zlo - z0:zhi - z0, ylo - y0:yhi - y0]
            full = new
            psi64, _, _ = oracle.view_update_whole(psi64, views[v], 0.0, dtype=np.float64)
            assert oracle.rel_l2(full, psi64) < 4e-6
            for d, dec, (ylo, yhi, y0, y1), (zlo, zhi, z0, z1) in shards:
                assert d.lib.dll.mvd_set_psi(d._ctx, np.ascontiguousarray(full[z0:z1, y0:y1]).ctypes.data_as(m._F)) == 0
    for d, *_ in shards:
        d.close()
