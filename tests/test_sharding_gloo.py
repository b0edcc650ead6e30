"""N > 1 path on CPU: two processes (gloo) each own a z-slab, run the view updates through the host-emulated kernel bodies and
exchange psi halos with mvrecon_b200.sharding.exchange_halos -- the same function bench.py uses with NCCL on the GPUs."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DIMS, VIEWS = (48, 20, 24), 2
KW = dict(psf_size_xyz=(5, 3, 5), psf_sigma_xyz=(1.0, 0.8, 1.4), bead_density=512)


def _worker(rank, world, port, lib_path, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import torch.distributed as dist
    import mvdecon_oracle as o
    import mvrecon_b200 as m
    from mvrecon_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = m.Lib(lib_path)
    ds = o.make_synthetic(DIMS, VIEWS, seed=5, **KW)
    views, psi0, avg = o.make_oracle_views(ds, o.EFFICIENT_BAYESIAN)
    nz, ny, nx = DIMS
    H = KW["psf_size_xyz"][2] - 1
    lo, hi = sharding.slab_range(nz, world, rank)
    z0, z1 = sharding.extended_range(lo, hi, nz, H)
    loc = [m.DeconView(ds.images[v][z0:z1], ds.weights[v][z0:z1], ds.psfs[v], m.PSFTYPE.EFFICIENT_BAYESIAN) for v in range(VIEWS)]
    dv = m.DeconViews(loc, shard=(lo, hi, z0, z1 - z0), global_dims_zyx=DIMS, library=lib)
    assert dv.halo_planes() == ((0 if rank == 0 else H), (0 if rank == world - 1 else H))
    dec = m.MultiViewDeconvolutionSeq(dv, 0, m.PsiInitFromRAI(psi0[z0:z1], [v.max_intensity for v in views]))
    plane = ny * nx
    for it in range(2):
        for v in range(VIEWS):
            dv.enqueue_view_update(v)
            dv.synchronize()
            buf = torch.from_numpy(dec.getPSI().reshape(-1).copy())
            sharding.exchange_halos(buf, plane, lo, hi, z0, H, rank, world, dist)
            assert lib.dll.mvd_set_psi(dv._ctx, buf.numpy().ctypes.data_as(m._F)) == 0
    np.save(os.path.join(out_dir, f"slab{rank}.npy"), dec.getPSI()[lo - z0:hi - z0])
    dv.close()
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_z_sharding_with_gloo_halo_exchange(hostemu_lib, oracle, tmp_path):
    import torch.multiprocessing as mp
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.start_processes(_worker, args=(world, port, hostemu_lib.path, str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    ds = oracle.make_synthetic(DIMS, VIEWS, seed=5, **KW)
    views, psi0, avg = oracle.make_oracle_views(ds, oracle.EFFICIENT_BAYESIAN)
    ref, _ = oracle.run_iterations_seq(psi0, views, 2, 0.0, dtype=np.float64)
    got = np.concatenate([np.load(tmp_path / f"slab{r}.npy") for r in range(world)], axis=0)
    assert got.shape == ref.shape
    assert oracle.rel_l2(got, ref) < 4e-6


def test_slab_ranges_cover_the_volume():
    sys.path.insert(0, ROOT)
    from mvrecon_b200 import sharding
    for nz, world in [(512, 8), (100, 3), (7, 7)]:
        r = [sharding.slab_range(nz, world, k) for k in range(world)]
        assert r[0][0] == 0 and r[-1][1] == nz and all(r[i][1] == r[i + 1][0] for i in range(world - 1))
    assert sharding.extended_range(64, 128, 512, 24) == (40, 152)
    assert sharding.extended_range(0, 64, 512, 24) == (0, 88)


def _worker2d(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from mvrecon_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nz, ny, nx, hy, hz, py, pz = 24, 28, 5, 3, 2, 2, 2
    ry, rz = rank // pz, rank % pz
    rank_of = lambda a, b: a * pz + b
    glob = torch.arange(nz * ny * nx, dtype=torch.float32).reshape(nz, ny, nx)
    ylo, yhi = sharding.slab_range(ny, py, ry)
    zlo, zhi = sharding.slab_range(nz, pz, rz)
    y0, y1 = sharding.extended_range(ylo, yhi, ny, hy)
    z0, z1 = sharding.extended_range(zlo, zhi, nz, hz)
    loc = torch.full((z1 - z0, y1 - y0, nx), -1.0)
    loc[zlo - z0:zhi - z0, ylo - y0:yhi - y0] = glob[zlo:zhi, ylo:yhi]          # only the owned box is known
    sharding.exchange_halos_2d(loc, (ylo, yhi), (y0, y1 - y0), (zlo, zhi), (z0, z1 - z0), hy, hz, ry, rz, py, pz, rank_of, dist)
    ok = bool(torch.equal(loc, glob[z0:z1, y0:y1]))                              # halos AND corners filled from the neighbours
    np.save(os.path.join(out_dir, f"ok{rank}.npy"), np.array([ok]))
    dist.barrier()
    dist.destroy_process_group()


def test_2d_halo_exchange_fills_halos_and_corners(tmp_path):
    import torch.multiprocessing as mp
    world = 4
    port = 31500 + (os.getpid() % 2000)
    mp.start_processes(_worker2d, args=(world, port, str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    assert all(bool(np.load(tmp_path / f"ok{r}.npy")[0]) for r in range(world))


def test_process_grid_prefers_y_first():
    sys.path.insert(0, ROOT)
    from mvrecon_b200 import sharding
    lengths = [32, 64, 120, 128, 180, 192, 256, 270, 288, 300, 320, 360, 512, 540, 576, 1024, 1080, 1152]
    assert sharding.grid_for(1, 1024, 512, 9, 12, lengths) == (1, 1)
    assert sharding.grid_for(2, 1024, 512, 9, 12, lengths) == (2, 1)          # the single-GPU plan already splits y in two tiles
    py, pz = sharding.grid_for(8, 1024, 512, 9, 12, lengths)
    assert py * pz == 8 and py >= 2
    assert sharding.axis_cost(1024, 0, 1024, 9, lengths) == 1080 and sharding.axis_cost(512, 0, 512, 12, lengths) == 540


# ---- exchange callback (mvd_set_exchange_callback) and exchange scheme 1 (psi + quotient exchange) ---------------------------------
DIMS_CB = (40, 28, 24)
KW_CB = dict(psf_size_xyz=(5, 5, 5), psf_sigma_xyz=(1.0, 1.1, 1.3), bead_density=512)


def _worker_cb(rank, world, port, lib_path, out_dir, axis, scheme, dims=None, max_len=0):
    os.environ["MVD_CHECK_RECTS"] = "1"   # the rectangle form of every filtered x launch must equal its box-filter form
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch.distributed as dist
    import mvdecon_oracle as o
    import mvrecon_b200 as m
    from mvrecon_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = m.Lib(lib_path)
    DIMS_CB = dims or globals()["DIMS_CB"]
    ds = o.make_synthetic(DIMS_CB, VIEWS, seed=6, **KW_CB)
    views, psi0, avg = o.make_oracle_views(ds, o.EFFICIENT_BAYESIAN)
    n = DIMS_CB[0] if axis == "z" else DIMS_CB[1]
    reach = 2                                                   # (5 - 1) / 2 for kernel1 and kernel2
    H = reach if scheme == 1 else 2 * reach
    lo, hi = sharding.slab_range(n, world, rank)
    a0, a1 = sharding.extended_range(lo, hi, n, H)
    sl = (slice(a0, a1),) if axis == "z" else (slice(None), slice(a0, a1))
    loc = [m.DeconView(np.ascontiguousarray(ds.images[v][sl]), np.ascontiguousarray(ds.weights[v][sl]), ds.psfs[v],
                       m.PSFTYPE.EFFICIENT_BAYESIAN) for v in range(VIEWS)]
    kw = {"shard": (lo, hi, a0, a1 - a0)} if axis == "z" else {"shard_y": (lo, hi, a0, a1 - a0)}
    dv = m.DeconViews(loc, global_dims_zyx=DIMS_CB, library=lib, exchange_scheme=scheme, max_fft_len=max_len, **kw)
    ntiles = dv.tile_info()["num_tiles"]
    assert (ntiles > 1) == (max_len > 0)
    want = ((0 if rank == 0 else H), (0 if rank == world - 1 else H))
    assert (dv.halo_planes() if axis == "z" else dv.halo_rows()) == want
    py, pz = (1, world) if axis == "z" else (world, 1)
    ry, rz = (0, rank) if axis == "z" else (rank, 0)
    calls = []
    inner = sharding.host_exchange_callback(ry, rz, py, pz, lambda a, b: a * pz + b, dist)

    def cb(which, box):
        calls.append(which)
        inner(which, box)

    dec = m.MultiViewDeconvolutionSeq(dv, 2, m.PsiInitFromRAI(np.ascontiguousarray(psi0[sl]), [v.max_intensity for v in views]))
    if scheme == 1:
        with pytest.raises(m.MvdError, match="exchange scheme 1"):
            dv.enqueue_view_update(0)
    dv.set_exchange_callback(cb)
    dec.runIterations()                                         # the library calls back for every exchange: no host-side loop
    assert calls == ([0] * 4 if scheme == 0 else ([1] * ntiles + [0]) * 4)       # scheme 1: the quotient spectrum of every x tile, then psi
    own = (slice(lo - a0, hi - a0),) if axis == "z" else (slice(None), slice(lo - a0, hi - a0))
    np.save(os.path.join(out_dir, f"part{rank}.npy"), dec.getPSI()[own])
    dv.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("axis,scheme", [("z", 0), ("z", 1), ("y", 1)])
def test_exchange_callback_and_two_exchange_scheme(hostemu_lib, oracle, tmp_path, axis, scheme):
    import torch.multiprocessing as mp
    world = 2
    port = 31500 + (os.getpid() % 2000) + 7 * scheme + (3 if axis == "y" else 0)
    mp.start_processes(_worker_cb, args=(world, port, hostemu_lib.path, str(tmp_path), axis, scheme), nprocs=world, join=True,
                       start_method="spawn")
    ds = oracle.make_synthetic(DIMS_CB, VIEWS, seed=6, **KW_CB)
    views, psi0, avg = oracle.make_oracle_views(ds, oracle.EFFICIENT_BAYESIAN)
    ref, _ = oracle.run_iterations_seq(psi0, views, 2, 0.0, dtype=np.float64)
    got = np.concatenate([np.load(tmp_path / f"part{r}.npy") for r in range(world)], axis=0 if axis == "z" else 1)
    assert got.shape == ref.shape
    assert oracle.rel_l2(got, ref) < 4e-6


def test_three_ranks_along_y_with_an_interior_rank(hostemu_lib, oracle, tmp_path):
    """N = 4 on the GPU box is a 4 x 1 grid: ranks in the middle have a neighbour on both y sides, so the boundary-first quotient pass runs
    with two interior sides and both z sides at volume faces.  Three gloo ranks reproduce that case."""
    import torch.multiprocessing as mp
    world, dims = 3, (40, 42, 24)
    port = 32300 + (os.getpid() % 2000)
    mp.start_processes(_worker_cb, args=(world, port, hostemu_lib.path, str(tmp_path), "y", 1, dims, 0), nprocs=world, join=True,
                       start_method="spawn")
    ds = oracle.make_synthetic(dims, VIEWS, seed=6, **KW_CB)
    views, psi0, avg = oracle.make_oracle_views(ds, oracle.EFFICIENT_BAYESIAN)
    ref, _ = oracle.run_iterations_seq(psi0, views, 2, 0.0, dtype=np.float64)
    got = np.concatenate([np.load(tmp_path / f"part{r}.npy") for r in range(world)], axis=1)
    assert got.shape == ref.shape
    assert oracle.rel_l2(got, ref) < 4e-6


def test_two_exchange_scheme_with_several_x_tiles(hostemu_lib, oracle, tmp_path):
    """exchange scheme 1 on a box that needs two FFT tiles along x (c4: 2048 + margins > one 2160-sample tile at equal cost): every tile
    exchanges its own quotient spectrum; y / z still fit one tile."""
    import torch.multiprocessing as mp
    world, dims = 2, (40, 28, 100)
    port = 31900 + (os.getpid() % 2000)
    mp.start_processes(_worker_cb, args=(world, port, hostemu_lib.path, str(tmp_path), "z", 1, dims, 32), nprocs=world, join=True,
                       start_method="spawn")
    ds = oracle.make_synthetic(dims, VIEWS, seed=6, **KW_CB)
    views, psi0, avg = oracle.make_oracle_views(ds, oracle.EFFICIENT_BAYESIAN)
    ref, _ = oracle.run_iterations_seq(psi0, views, 2, 0.0, dtype=np.float64)
    got = np.concatenate([np.load(tmp_path / f"part{r}.npy") for r in range(world)], axis=0)
    assert got.shape == ref.shape
    assert oracle.rel_l2(got, ref) < 4e-6


# ---- 2-d (y x z) grid of four boxes: corners must be right under both exchange schemes ----------------------------------------------
DIMS_2D = (40, 44, 24)


def _worker_grid(rank, world, port, lib_path, out_dir, scheme):
    os.environ["MVD_SPLIT_P1"] = "1"      # also cover the (opt-in) split of the forward pass around the psi exchange
    os.environ["MVD_CHECK_RECTS"] = "1"
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch.distributed as dist
    import mvdecon_oracle as o
    import mvrecon_b200 as m
    from mvrecon_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = m.Lib(lib_path)
    ds = o.make_synthetic(DIMS_2D, VIEWS, seed=8, **KW_CB)
    views, psi0, avg = o.make_oracle_views(ds, o.EFFICIENT_BAYESIAN)
    nz, ny, nx = DIMS_2D
    py, pz = 2, 2
    ry, rz = rank // pz, rank % pz
    H = 2 if scheme == 1 else 4
    ylo, yhi = sharding.slab_range(ny, py, ry)
    zlo, zhi = sharding.slab_range(nz, pz, rz)
    y0, y1 = sharding.extended_range(ylo, yhi, ny, H)
    z0, z1 = sharding.extended_range(zlo, zhi, nz, H)
    cut = lambda a: np.ascontiguousarray(a[z0:z1, y0:y1])
    loc = [m.DeconView(cut(ds.images[v]), cut(ds.weights[v]), ds.psfs[v], m.PSFTYPE.EFFICIENT_BAYESIAN) for v in range(VIEWS)]
    dv = m.DeconViews(loc, global_dims_zyx=DIMS_2D, library=lib, exchange_scheme=scheme,
                      shard=(zlo, zhi, z0, z1 - z0), shard_y=(ylo, yhi, y0, y1 - y0))
    dv.set_exchange_callback(sharding.host_exchange_callback(ry, rz, py, pz, lambda a, b: a * pz + b, dist))
    dec = m.MultiViewDeconvolutionSeq(dv, 2, m.PsiInitFromRAI(cut(psi0), [v.max_intensity for v in views]))
    dec.runIterations()
    np.save(os.path.join(out_dir, f"box{rank}.npy"), dec.getPSI()[zlo - z0:zhi - z0, ylo - y0:yhi - y0])
    dv.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("scheme", [0, 1])
def test_four_boxes_2d_grid(hostemu_lib, oracle, tmp_path, scheme):
    import torch.multiprocessing as mp
    from mvrecon_b200 import sharding
    world = 4
    port = 33500 + (os.getpid() % 2000) + 13 * scheme
    mp.start_processes(_worker_grid, args=(world, port, hostemu_lib.path, str(tmp_path), scheme), nprocs=world, join=True, start_method="spawn")
    ds = oracle.make_synthetic(DIMS_2D, VIEWS, seed=8, **KW_CB)
    views, psi0, avg = oracle.make_oracle_views(ds, oracle.EFFICIENT_BAYESIAN)
    ref, _ = oracle.run_iterations_seq(psi0, views, 2, 0.0, dtype=np.float64)
    got = np.empty_like(ref, dtype=np.float32)
    nz, ny, nx = DIMS_2D
    for r in range(world):
        ry, rz = r // 2, r % 2
        ylo, yhi = sharding.slab_range(ny, 2, ry)
        zlo, zhi = sharding.slab_range(nz, 2, rz)
        got[zlo:zhi, ylo:yhi] = np.load(tmp_path / f"box{r}.npy")
    assert oracle.rel_l2(got, ref) < 4e-6


# ---- PsiInit, per-view maxima and IterationStatistics on a 2 x 2 (y x z) grid: global quantities through the reduce callback -------
def _worker_psiinit(rank, world, port, lib_path, out_dir, kind):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch.distributed as dist
    import mvdecon_oracle as o
    import mvrecon_b200 as m
    from mvrecon_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = m.Lib(lib_path)
    ds = o.make_synthetic(DIMS_2D, VIEWS, seed=8, **KW_CB)
    nz, ny, nx = DIMS_2D
    py, pz = 2, 2
    ry, rz = rank // pz, rank % pz
    H = 2                                                       # exchange scheme 1: max(k1/2, k2/2)
    ylo, yhi = sharding.slab_range(ny, py, ry)
    zlo, zhi = sharding.slab_range(nz, pz, rz)
    y0, y1 = sharding.extended_range(ylo, yhi, ny, H)
    z0, z1 = sharding.extended_range(zlo, zhi, nz, H)
    cut = lambda a: np.ascontiguousarray(a[z0:z1, y0:y1])
    loc = [m.DeconView(cut(ds.images[v]), cut(ds.weights[v]), ds.psfs[v], m.PSFTYPE.EFFICIENT_BAYESIAN) for v in range(VIEWS)]
    dv = m.DeconViews(loc, global_dims_zyx=DIMS_2D, library=lib, exchange_scheme=1,
                      shard=(zlo, zhi, z0, z1 - z0), shard_y=(ylo, yhi, y0, y1 - y0))
    init = {"fused": m.PsiInitBlurredFused(1.5), "avg": m.PsiInitAvgPrecise(), "approx": m.PsiInitAvgApprox()}[kind]
    with pytest.raises(m.MvdError, match="sharded"):            # global statistics need the reduce plumbing first
        init.runInitialization(dv)
    dv.set_exchange_callback(sharding.host_exchange_callback(ry, rz, py, pz, lambda a, b: a * pz + b, dist))
    dv.set_reduce_callback(sharding.host_reduce_callback(dist))
    dec = m.MultiViewDeconvolutionSeq(dv, 1, init)
    assert dec.initWasSuccessful()
    np.save(os.path.join(out_dir, f"psi0_{rank}.npy"), dec.getPSI())           # the whole local array: halos must hold the neighbours' values
    dec.runIterations()
    np.save(os.path.join(out_dir, f"box{rank}.npy"), dec.getPSI()[zlo - z0:zhi - z0, ylo - y0:yhi - y0])
    np.save(os.path.join(out_dir, f"meta{rank}.npy"), np.array([init.getAvg()] + list(init.getMax()) +
                                                               [x for s in dec.stats[0] for x in (s.sumChange, s.maxChange)]))
    dv.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["fused", "avg", "approx"])
def test_psi_init_and_statistics_on_a_2d_grid(hostemu_lib, oracle, tmp_path, kind):
    """PsiInitBlurredFused / AvgPrecise / AvgApprox on y x z sharded contexts (PsiInitBlurredFused.java:76-127, MultiViewDeconvolution.java:115-135):
    psi0, avg and max[] equal the whole-volume values on every rank; the statistics of a view update are summed / maxed over the boxes
    (MultiViewDeconvolutionSeq.java:165-176)."""
    import torch.multiprocessing as mp
    from mvrecon_b200 import sharding
    world = 4
    port = 35500 + (os.getpid() % 2000) + {"fused": 0, "avg": 17, "approx": 29}[kind]
    mp.start_processes(_worker_psiinit, args=(world, port, hostemu_lib.path, str(tmp_path), kind), nprocs=world, join=True, start_method="spawn")
    ds = oracle.make_synthetic(DIMS_2D, VIEWS, seed=8, **KW_CB)
    if kind == "fused":
        psi0, mx, avg = oracle.psi_init_blurred_fused(ds.images, ds.weights, 1.5)
    elif kind == "avg":
        psi0, mx, avg = oracle.psi_init_avg_precise(ds.images)
    else:
        psi0, mx, avg = oracle.psi_init_avg_approx(ds.images)
    k1, k2 = oracle.derive_kernels(ds.psfs, oracle.EFFICIENT_BAYESIAN)
    views = [oracle.OracleView(ds.images[v], ds.weights[v], k1[v], k2[v], float(mx[v])) for v in range(VIEWS)]
    ref, st = oracle.run_iterations_seq(psi0, views, 1, 0.0, dtype=np.float64)
    nz, ny, nx = DIMS_2D
    got = np.empty_like(ref, dtype=np.float32)
    for r in range(world):
        ry, rz = r // 2, r % 2
        ylo, yhi = sharding.slab_range(ny, 2, ry)
        zlo, zhi = sharding.slab_range(nz, 2, rz)
        y0, y1 = sharding.extended_range(ylo, yhi, ny, 2)
        z0, z1 = sharding.extended_range(zlo, zhi, nz, 2)
        loc0 = np.load(tmp_path / f"psi0_{r}.npy")
        assert oracle.rel_l2(loc0, psi0[z0:z1, y0:y1]) < 1e-6, (r, kind)
        got[zlo:zhi, ylo:yhi] = np.load(tmp_path / f"box{r}.npy")
        meta = np.load(tmp_path / f"meta{r}.npy")
        assert abs(meta[0] - avg) <= 1e-9 * abs(avg)
        assert np.array_equal(meta[1:1 + VIEWS].astype(np.float32), mx)
        for v in range(VIEWS):
            s_ref, m_ref = st[v][2], st[v][3]
            assert abs(meta[1 + VIEWS + 2 * v] - s_ref) <= 1e-4 * abs(s_ref) + 0.5
            assert abs(meta[2 + VIEWS + 2 * v] - m_ref) <= 1e-3 * abs(m_ref) + 1e-3
    assert oracle.rel_l2(got, ref) < 4e-6
