#!/usr/bin/env python3
"""Generates tests/golden/small_case.npz from the float64 oracle.

The reference ships no golden vectors for this path (no JUnit, no fixtures -> "parity unpinned"), and
it cannot be executed in this image (Java).  The vectors below therefore pin the *oracle* (and through it the CUDA path)
against accidental drift: seeded inputs -> kernel1/kernel2 per PSFTYPE, psi after every view update of iterations 1-2,
psi after iteration 5, and the signed statistics.  Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import mvdecon_oracle as o  # noqa: E402

DIMS, VIEWS, SEED = (33, 36, 40), 3, 1
KW = dict(psf_size_xyz=(7, 5, 7), psf_sigma_xyz=(1.2, 1.0, 2.0), bead_density=512)
LAMBDA = 0.006
QUIRK_T = 8      # Threads.numThreads() of the pinned reference run (AdjustInput.sumImg double-counts portion 0 of max(T, size/64^3) portions)


def main():
    ds = o.make_synthetic(DIMS, VIEWS, seed=SEED, **KW)
    out = {"dims_zyx": np.array(DIMS), "lambda": np.array(LAMBDA), "quirk_threads": np.array(QUIRK_T)}
    # inputs are stored too (float16-exact? no: float32, small) so the fixture does not depend on the generator staying fixed
    for v in range(VIEWS):
        out[f"img{v}"] = ds.images[v]
        out[f"weight{v}"] = ds.weights[v]
        out[f"psf{v}"] = ds.psfs[v]
    for ptype in range(4):
        k1, k2 = o.derive_kernels(ds.psfs, ptype, quirk_threads=QUIRK_T, dtype=np.float64)
        for v in range(VIEWS):
            out[f"k1_t{ptype}_v{v}"] = k1[v]
            out[f"k2_t{ptype}_v{v}"] = k2[v]
    views, psi0, avg = o.make_oracle_views(ds, o.EFFICIENT_BAYESIAN, quirk_threads=QUIRK_T)
    out["psi0"] = psi0
    out["max"] = np.array([v.max_intensity for v in views], dtype=np.float32)
    stats = []

    def cb(it, v, psi, s, m):
        stats.append((it, v, s, m))
        if it < 2:
            out[f"psi_it{it}_v{v}"] = psi.copy()

    psi, _ = o.run_iterations_seq(psi0, views, 5, LAMBDA, dtype=np.float64, callback=cb)
    out["psi_it4"] = psi
    out["stats"] = np.array(stats, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "small_case.npz"), **out)
    print("wrote small_case.npz", sum(v.nbytes for v in out.values()) // 1024, "KiB uncompressed")


if __name__ == "__main__":
    main()
