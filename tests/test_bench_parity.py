"""bench.py's in-run parity check (crop + halo against the float64 oracle) validated on the CPU: the region update must reproduce the
whole-volume oracle on the crop core, the region gather must reassemble a volume from the ranks' boxes, and the synthetic generator of
bench.py must agree with the oracle's."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_region_update_equals_whole_volume_on_the_core(oracle):
    import bench
    W = dict(dims=(60, 64, 72), views=3, psf=(5, 5, 7), sigma=(1.0, 1.0, 1.6), lam=0.006, ptype=2, tilt=None, iters=1)
    ds = oracle.make_synthetic(W["dims"], W["views"], seed=bench.SEED, psf_size_xyz=W["psf"], psf_sigma_xyz=W["sigma"], bead_density=2048)
    for v in range(W["views"]):                     # bench.py's generator: same PSFs, same coverage boxes as the oracle's
        assert np.array_equal(bench.synth_psf(v, W["views"], W["psf"], W["sigma"]), ds.psfs[v])
        mn, mx = bench.coverage_box(W["dims"], v)
        assert (tuple(mn), tuple(mx)) == ds.boxes[v]
    assert np.array_equal(bench.truth_box(W["dims"], 10, 50, 5, 40, bench.SEED), oracle.synth_truth(W["dims"], bench.SEED)[5:40, 10:50])
    views, psi0, avg = oracle.make_oracle_views(ds, W["ptype"])
    mx = [v.max_intensity for v in views]
    nv = 2
    full = psi0
    for v in range(nv):
        full, _, _ = oracle.view_update_whole(full, views[v], W["lam"], dtype=np.float64)
    region, core = bench.parity_region(W, nv, 32, 30, core=8)
    rs = tuple(slice(lo, hi) for lo, hi in region)
    assert region[0] == (30 - 4 - 12, 30 + 4 + 12) and region[2][0] > 0          # interior cut faces in z and x: only the core is trusted
    got = bench.oracle_region_update(W, ds.psfs, region, psi0[rs], [im[rs] for im in ds.images], mx, nv)
    cs = tuple(slice(c[0] - r[0], c[1] - r[0]) for c, r in zip(core, region))
    gs = tuple(slice(c[0], c[1]) for c in core)
    assert oracle.rel_l2(got[cs], full[gs]) < 1e-12
    # a region clipped by the volume (true faces) is exact everywhere except near its interior cuts
    region2, core2 = bench.parity_region(W, nv, 2, 2, core=8)
    assert region2[0][0] == 0 and region2[1][0] == 0
    rs2 = tuple(slice(lo, hi) for lo, hi in region2)
    got2 = bench.oracle_region_update(W, ds.psfs, region2, psi0[rs2], [im[rs2] for im in ds.images], mx, nv)
    cs2 = tuple(slice(c[0] - r[0], c[1] - r[0]) for c, r in zip(core2, region2))
    gs2 = tuple(slice(c[0], c[1]) for c in core2)
    assert oracle.rel_l2(got2[cs2], full[gs2]) < 1e-12


def test_gather_region_reassembles_the_boxes():
    import torch
    import bench
    rng = np.random.default_rng(4)
    vol = torch.from_numpy(rng.random((20, 24, 10)).astype(np.float32))
    region = [(6, 17), (5, 21), (2, 9)]
    parts = []
    for (zlo, zhi) in ((0, 9), (9, 20)):
        for (ylo, yhi) in ((0, 12), (12, 24)):
            z0, z1, y0, y1 = max(0, zlo - 2), min(20, zhi + 2), max(0, ylo - 3), min(24, yhi + 3)
            local = vol[z0:z1, y0:y1].clone()
            local[:zlo - z0] = -7; local[:, :ylo - y0] = -7           # halos hold junk: only own parts may be used
            if zhi - z0 < local.shape[0]:
                local[zhi - z0:] = -7
            if yhi - y0 < local.shape[1]:
                local[:, yhi - y0:] = -7
            parts.append(bench.gather_region(torch, None, local, ((zlo, zhi), (ylo, yhi)), (z0, y0), region))
    total = sum(parts)
    assert torch.equal(total, vol[6:17, 5:21, 2:9])
