"""Two contexts of the host emulator in one process (two threads), joined by mvd_set_exchange_callback with a plain numpy copy between
the two boxes: exercises the library's box / halo-width bookkeeping for EVEN kernel extents, where the reach below and above a voxel
differ (hy_lo != hy_hi), under both exchange schemes and both sharding axes."""
import ctypes as C
import threading

import numpy as np
import pytest

DIMS = (26, 28, 20)            # z, y, x
PSF_XYZ = (3, 4, 6)            # even in y and z


def _reach(k):                 # csrc/engine.h reach_of
    return k - 1 - k // 2, k // 2


@pytest.mark.parametrize("axis,scheme", [("z", 0), ("z", 1), ("y", 0), ("y", 1)])
def test_even_kernels_two_boxes(hostemu_lib, oracle, axis, scheme):
    import mvrecon_b200 as m
    rng = np.random.default_rng(31)
    V = 2
    kd = (PSF_XYZ[2], PSF_XYZ[1], PSF_XYZ[0])
    psfs = [(0.05 + rng.random(kd)).astype(np.float32) for _ in range(V)]
    imgs = [(1.0 + 50.0 * rng.random(DIMS)).astype(np.float32) for _ in range(V)]
    ws = [(rng.random(DIMS) / V).astype(np.float32) for _ in range(V)]
    psi0 = (5.0 + 20.0 * rng.random(DIMS)).astype(np.float32)
    mx = [float(im.max()) for im in imgs]
    k1, k2 = oracle.derive_kernels(psfs, oracle.INDEPENDENT)
    views = [oracle.OracleView(imgs[v], ws[v], k1[v], k2[v], mx[v]) for v in range(V)]
    ref = psi0
    for it in range(2):
        for v in range(V):
            ref, _, _ = oracle.view_update_whole(ref, views[v], 0.0, dtype=np.float64)

    ax = 0 if axis == "z" else 1
    n = DIMS[ax]
    lo_r, hi_r = _reach(PSF_XYZ[2] if axis == "z" else PSF_XYZ[1])
    need_lo, need_hi = (lo_r, hi_r) if scheme == 1 else (2 * lo_r, 2 * hi_r)          # kernel1 and kernel2 have the same extents
    cut = n // 2 + 1
    own = [(0, cut), (cut, n)]
    ext = [(0, min(n, cut + need_hi)), (max(0, cut - need_lo), n)]
    boxes, res, err = {}, {}, []
    bar = threading.Barrier(2)

    def work(r):
        try:
            (lo, hi), (a0, a1) = own[r], ext[r]
            sl = (slice(a0, a1),) if axis == "z" else (slice(None), slice(a0, a1))
            loc = [m.DeconView(np.ascontiguousarray(imgs[v][sl]), np.ascontiguousarray(ws[v][sl]), psfs[v], m.PSFTYPE.INDEPENDENT) for v in range(V)]
            kw = {"shard": (lo, hi, a0, a1 - a0)} if axis == "z" else {"shard_y": (lo, hi, a0, a1 - a0)}
            dv = m.DeconViews(loc, global_dims_zyx=DIMS, library=hostemu_lib, exchange_scheme=scheme, **kw)
            want = (0, need_hi) if r == 0 else (need_lo, 0)
            assert (dv.halo_planes() if axis == "z" else dv.halo_rows()) == want

            def cb(which, box):
                arr = np.ctypeslib.as_array(box.base, shape=(box.nplanes, box.nrows, box.row_floats))
                h_lo, h_hi = (box.hz_lo, box.hz_hi) if axis == "z" else (box.hy_lo, box.hy_hi)
                o0, o1 = (box.z0, box.z1) if axis == "z" else (box.y0, box.y1)
                boxes[r] = (arr, o0, o1)
                bar.wait()
                other, p0, p1 = boxes[1 - r]
                take = (lambda a, s: a[s]) if axis == "z" else (lambda a, s: a[:, s])
                if r == 0:       # upper neighbour's first h_hi own rows / planes -> my upper halo
                    dst, src = slice(o1, o1 + h_hi), slice(p0, p0 + h_hi)
                else:            # lower neighbour's last h_lo own rows / planes -> my lower halo
                    dst, src = slice(o0 - h_lo, o0), slice(p1 - h_lo, p1)
                if axis == "z":
                    arr[dst] = other[src]
                else:
                    arr[:, dst] = take(other, src)
                bar.wait()

            dv.set_exchange_callback(cb)
            dec = m.MultiViewDeconvolutionSeq(dv, 2, m.PsiInitFromRAI(np.ascontiguousarray(psi0[sl]), mx))
            dec.runIterations()
            osl = (slice(lo - a0, hi - a0),) if axis == "z" else (slice(None), slice(lo - a0, hi - a0))
            res[r] = dec.getPSI()[osl].copy()
            dv.close()
        except Exception as e:  # noqa: BLE001
            err.append(e)
            bar.abort()

    ts = [threading.Thread(target=work, args=(r,), daemon=True) for r in (0, 1)]
    [t.start() for t in ts]
    [t.join(timeout=120) for t in ts]
    assert not err, err
    got = np.concatenate([res[0], res[1]], axis=ax)
    assert oracle.rel_l2(got, ref) < 4e-6
