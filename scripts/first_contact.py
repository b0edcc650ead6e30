#!/usr/bin/env python3
"""First GPU contact: small parity checks against the oracle + raw timing at c2/c3 sizes. Development helper."""
import json
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import mvrecon_b200 as m  # noqa: E402
import mvdecon_oracle as o  # noqa: E402

OUT = {}


def stage(name):
    def deco(fn):
        t = time.time()
        try:
            OUT[name] = fn()
        except Exception as e:  # noqa: BLE001
            OUT[name] = {"error": repr(e), "tb": traceback.format_exc()[-1500:]}
        OUT[name + "_s"] = round(time.time() - t, 2)
        print(name, json.dumps(OUT[name], default=str)[:2000], flush=True)
    return deco


L = m.lib()
print("devices", L.getNumDevicesCUDA(), L.getNameDeviceCUDA(0), flush=True)


@stage("conv_small")
def _():
    rng = np.random.default_rng(0)
    res = {}
    for shape, ks in [((9, 10, 11), (3, 5, 3)), ((40, 50, 70), (7, 5, 9)), ((130, 100, 90), (25, 19, 25))]:
        img = rng.random(shape).astype(np.float32)
        k = rng.random(ks).astype(np.float32)
        for ext in ("mirror", "zero", "const"):
            a = L.convolve(img, k, ext, ext_value=1.0)
            b = o.fft_convolve(img, k, ext, const=1.0, dtype=np.float64)
            res[f"{shape}_{ext}"] = o.rel_l2(a, b)
    return res


@stage("legacy_circular")
def _():
    rng = np.random.default_rng(1)
    res = {}
    for shape in [(32, 32, 32), (64, 32, 128)]:
        img = rng.random(shape).astype(np.float32)
        k = rng.random((5, 7, 9)).astype(np.float32)
        ref = o.circular_convolve(img, k, dtype=np.float64)
        im2 = img.copy()
        L.convolution3DfftCUDAInPlace(im2, k, 0)
        res[str(shape)] = o.rel_l2(im2, ref)
    return res


@stage("loop_small")
def _():
    res = {}
    ds = o.make_synthetic((33, 36, 40), 3, seed=1, psf_size_xyz=(7, 5, 7), psf_sigma_xyz=(1.2, 1.0, 2.0), bead_density=512)
    for ptype in (o.EFFICIENT_BAYESIAN, o.INDEPENDENT):
        views, psi0, avg = o.make_oracle_views(ds, ptype)
        dv = m.DeconViews([m.DeconView(ds.images[v], ds.weights[v], ds.psfs[v], m.PSFTYPE(ptype)) for v in range(3)], lambda_=0.006)
        dec = m.MultiViewDeconvolutionSeq(dv, 3, m.PsiInitFromRAI(psi0, [v.max_intensity for v in views]))
        dec.runIterations()
        psi = dec.getPSI()
        p64, st = o.run_iterations_seq(psi0, views, 3, 0.006, dtype=np.float64)
        res[o.PSFTYPE_NAMES[ptype]] = {"relL2": o.rel_l2(psi, p64), "maxabs": float(np.abs(psi - p64).max()),
                                       "stats_gpu": [dec.stats[0][0].sumChange, dec.stats[0][0].maxChange], "stats_ref": list(st[0][2:])}
        dv.close()
    return res


@stage("loop_multitile")
def _():
    # force several tiles per axis with a small max FFT length
    ds = o.make_synthetic((70, 90, 100), 2, seed=2, psf_size_xyz=(9, 7, 9), psf_sigma_xyz=(1.5, 1.2, 2.5), bead_density=2048)
    views, psi0, avg = o.make_oracle_views(ds, o.EFFICIENT_BAYESIAN)
    dv = m.DeconViews([m.DeconView(ds.images[v], ds.weights[v], ds.psfs[v], m.PSFTYPE.EFFICIENT_BAYESIAN) for v in range(2)], lambda_=0.0, max_fft_len=48)
    info = dv.tile_info()
    dec = m.MultiViewDeconvolutionSeq(dv, 2, m.PsiInitFromRAI(psi0, [v.max_intensity for v in views]))
    dec.runIterations()
    psi = dec.getPSI()
    p64, st = o.run_iterations_seq(psi0, views, 2, 0.0, dtype=np.float64)
    dv.close()
    return {"tiles": info, "relL2": o.rel_l2(psi, p64), "maxabs": float(np.abs(psi - p64).max())}


def timing(dims_zyx, nviews, psf_xyz, lam, iters, max_fft_len=0):
    rng = np.random.default_rng(3)
    nz, ny, nx = dims_zyx
    views = []
    psfs = [o.synth_psf(v, nviews, psf_xyz) for v in range(nviews)]
    base = (100.0 + 50.0 * rng.random(dims_zyx, dtype=np.float32)).astype(np.float32)
    w = np.full(dims_zyx, 1.0 / nviews, dtype=np.float32)
    for v in range(nviews):
        views.append(m.DeconView(base, w, psfs[v], m.PSFTYPE.EFFICIENT_BAYESIAN))
    t0 = time.time()
    dv = m.DeconViews(views, lambda_=lam, max_fft_len=max_fft_len)
    t_init = time.time() - t0
    info = dv.tile_info()
    dec = m.MultiViewDeconvolutionSeq(dv, 1, m.PsiInitFromRAI(base, [200.0] * nviews))
    dec.runIterations()            # warm-up iteration
    dec.numIterations = 1 + iters
    t0 = time.time()
    dec.runIterations()
    dt = time.time() - t0
    psi = dec.getPSI()
    dv.close()
    vox = nz * ny * nx
    thr = vox * nviews * iters / dt
    return {"tiles": info, "init_s": round(t_init, 2), "loop_s": round(dt, 4), "Gvox_view_it_per_s": round(thr / 1e9, 3),
            "roofline_frac_92B_6542": round(thr * 92 / 6542.7e9, 4), "psi_finite": bool(np.isfinite(psi).all()), "psi_mean": float(psi.mean())}


@stage("time_c1")
def _():
    return timing((128, 256, 256), 4, (25, 19, 25), 0.0, 3)


@stage("time_c2")
def _():
    return timing((256, 512, 512), 6, (25, 19, 25), 0.006, 3)


@stage("time_c3")
def _():
    return timing((512, 1024, 1024), 4, (25, 19, 25), 0.0, 3)


os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "first_contact.json"), "w") as f:
    json.dump(OUT, f, indent=1, default=str)
