#!/usr/bin/env python3
"""Per-pass CUDA-event timing of one workload (development helper): python scripts/prof_passes.py c3 [iters] [max_fft_len]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import mvrecon_b200 as m  # noqa: E402
import mvdecon_oracle as o  # noqa: E402

CFG = {
    "c1": ((128, 256, 256), 4, (25, 19, 25), 0.0),
    "c2": ((256, 512, 512), 6, (25, 19, 25), 0.006),
    "c3": ((512, 1024, 1024), 4, (25, 19, 25), 0.0),
}
# bytes moved per FFT-box voxel (real voxel) by each pass: P1..P9
BOX_BYTES = [8, 8, 12, 8, 12, 8, 12, 8, 16]


def main():
    if os.environ.get("MVD_LIB"):                 # development only: A/B a variant build of the library
        m._LIB = m.Lib(os.environ["MVD_LIB"])
    name = sys.argv[1] if len(sys.argv) > 1 else "c3"
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    max_len = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    dims, nviews, psf_xyz, lam = CFG[name]
    rng = np.random.default_rng(3)
    psfs = [o.synth_psf(v, nviews, psf_xyz) for v in range(nviews)]
    base = (100.0 + 50.0 * rng.random(dims, dtype=np.float32)).astype(np.float32)
    w = np.full(dims, 1.0 / nviews, dtype=np.float32)
    imgs = [base] * nviews
    ws = [w] * nviews
    if os.environ.get("MVD_REALISTIC"):           # coverage gaps like the bench data: zero image / zero weight slabs, graded weights
        imgs, ws = [], []
        for v in range(nviews):
            im = base.copy(); wv = w.copy()
            sl = [slice(None)] * 3
            ax = (v % 6) // 2
            cut = dims[ax] // 8
            sl[ax] = slice(0, cut) if v % 2 == 0 else slice(dims[ax] - cut, dims[ax])
            im[tuple(sl)] = 0; wv[tuple(sl)] = 0
            ramp = np.linspace(0, 1, dims[2], dtype=np.float32)
            wv *= ramp[None, None, :]
            imgs.append(im); ws.append(wv)
    dv = m.DeconViews([m.DeconView(imgs[v], ws[v], psfs[v], m.PSFTYPE.EFFICIENT_BAYESIAN) for v in range(nviews)], lambda_=lam, max_fft_len=max_len)
    info = dv.tile_info()
    dec = m.MultiViewDeconvolutionSeq(dv, 1, m.PsiInitFromRAI(base, [200.0] * nviews))
    dec.runIterations()
    dv.set_profiling(True)
    dec.numIterations = 1 + iters
    t0 = time.time()
    dec.runIterations()
    dt = time.time() - t0
    ms, n = dv.pass_times()
    vox = dims[0] * dims[1] * dims[2]
    tx, ty, tz = info["tile_dims_xyz"]
    box = tx * ty * tz
    out = {"config": name, "tiles": info, "loop_s": dt, "Gvvi_s": vox * nviews * iters / dt / 1e9, "passes": []}
    tot = sum(ms)
    for i in range(9):
        per = ms[i] / max(n[i], 1)
        out["passes"].append({"pass": f"P{i + 1}", "ms_per_launch": round(per, 4), "share": round(ms[i] / tot, 4),
                              "box_GBs": round(BOX_BYTES[i] * box / (per * 1e-3) / 1e9, 1) if per > 0 else None})
    out["sum_ms_per_view_update"] = tot / (iters * nviews)
    print(json.dumps(out, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"prof_passes_{name}.json"), "w") as f:
        json.dump(out, f, indent=1)
    dv.close()


if __name__ == "__main__":
    main()
