#!/usr/bin/env python3
"""Generates tests/golden/c1_case.npz: BASELINE config c1 in full (4 views, 256x256x128, PSF 25x19x25, EFFICIENT_BAYESIAN, lambda = 0,
10 iterations) from the float64 oracle with the reference's PSF normalisation for a pinned thread count (T = 8).

The volumes are too large to commit (32 MiB per snapshot), so the fixture holds: psi after iterations 1, 2 and 10 on a lattice of every
8th voxel per axis (16384 samples each), the float64 sum and sum of squares of those snapshots, the signed statistics of all 40 view
updates, the per-view maxima, the PsiInit average, and checksums of the seeded inputs (so a drifting generator is noticed).
The reference ships no golden vectors for this path and cannot run here (Java): this pins the oracle.
Run:  python tests/golden/make_golden_c1.py      (about a minute on 8 cores)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import mvdecon_oracle as o  # noqa: E402

DIMS, VIEWS, SEED, ITERS, QUIRK_T, STEP = (128, 256, 256), 4, 20261, 10, 8, 8


def inputs():
    ds = o.make_synthetic(DIMS, VIEWS, seed=SEED)                      # PSF 25x19x25, sigma (1.5, 1.5, 4.0), blending range 12
    views, psi0, avg = o.make_oracle_views(ds, o.EFFICIENT_BAYESIAN, quirk_threads=QUIRK_T)
    return ds, views, psi0, avg


def main():
    ds, views, psi0, avg = inputs()
    out = {"dims_zyx": np.array(DIMS), "quirk_threads": np.array(QUIRK_T), "lattice_step": np.array(STEP), "avg": np.array(avg),
           "max": np.array([v.max_intensity for v in views], dtype=np.float32),
           "img_sums": np.array([float(im.sum(dtype=np.float64)) for im in ds.images]),
           "weight_sums": np.array([float(w.sum(dtype=np.float64)) for w in ds.weights]),
           "psi0_sum": np.array(float(psi0.sum(dtype=np.float64))),
           "k1_sums": np.array([float(v.kernel1.sum(dtype=np.float64)) for v in views]),
           "k2_sums": np.array([float(v.kernel2.sum(dtype=np.float64)) for v in views])}
    stats = []
    psi = psi0
    for it in range(ITERS):
        for v in range(VIEWS):
            psi, s, m = o.view_update_whole(psi, views[v], 0.0, dtype=np.float64)
            stats.append((it, v, s, m))
        if it + 1 in (1, 2, ITERS):
            out[f"psi_it{it + 1}"] = psi[::STEP, ::STEP, ::STEP].astype(np.float32)
            p = psi.astype(np.float64)
            out[f"psi_it{it + 1}_moments"] = np.array([p.sum(), (p * p).sum()])
        print("iteration", it + 1, "sumChange", stats[-1][2], flush=True)
    out["stats"] = np.array(stats, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "c1_case.npz"), **out)
    print("wrote c1_case.npz", sum(v.nbytes for v in out.values()) // 1024, "KiB uncompressed")


if __name__ == "__main__":
    main()
