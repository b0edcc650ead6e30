#!/usr/bin/env python3
"""Generates tests/golden/prep_case.npz from the oracle: the step before the loop (view materialisation, PSF preparation, weight masks).
Like small_case.npz these vectors pin the *oracle* (parity is unpinned against the Java reference, which cannot run here) so that neither it
nor the device code can drift unnoticed.  Run:  python tests/golden/make_golden_prep.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import mvdecon_oracle as o  # noqa: E402

DIMS, BBOX_MIN = (18, 22, 26), (-3, 1, 2)


def models():
    out = []
    for th_deg, sz, tr in ((27.0, 1.7, (9.3, -2.25, 14.1)), (-14.0, 2.1, (4.2, 1.5, -3.75))):
        th = np.deg2rad(th_deg)
        fwd = np.array([[np.cos(th), 0.0, np.sin(th) * sz, tr[0]], [0.0, 1.0, 0.0, tr[1]], [-np.sin(th), 0.0, np.cos(th) * sz, tr[2]], [0, 0, 0, 1.0]])
        out.append((fwd[:3].ravel(), np.linalg.inv(fwd)[:3].ravel()))
    return out


def main():
    rng = np.random.default_rng(41)
    raws = [(rng.random(s) * 300).astype(np.float32) for s in ((12, 22, 24), (11, 20, 22))]
    psfs = [rng.random(s).astype(np.float32) for s in ((7, 9, 11), (9, 9, 9))]
    ms = models()
    fb = [((2.0, 1.0, 0.5), (12.0, 10.0, 6.0)), ((1.0, 1.0, 1.0), (8.0, 8.0, 4.0))]
    db = [((-3.0, -3.0, -1.0), (12.0, 10.0, 6.0)), ((0.0, 0.0, 0.0), (6.0, 6.0, 3.0))]
    out = {"dims_zyx": np.array(DIMS), "bbox_min_xyz": np.array(BBOX_MIN), "fusion_blending": np.array(fb, dtype=np.float32),
           "decon_blending": np.array(db, dtype=np.float32)}
    for j in range(2):
        out[f"raw{j}"] = raws[j]; out[f"psf{j}"] = psfs[j]; out[f"affine{j}"] = ms[j][0]; out[f"inv_affine{j}"] = ms[j][1]
        out[f"view_linear{j}"] = o.transform_view(raws[j], ms[j][1], BBOX_MIN, DIMS, 1)
        out[f"view_nearest{j}"] = o.transform_view(raws[j], ms[j][1], BBOX_MIN, DIMS, 0)
        out[f"psf_t{j}"] = o.psf_transform(psfs[j], *ms[j])
    img, w = o.fuse_group(raws, [m[1] for m in ms], BBOX_MIN, DIMS, 1, fb, db)
    out["fused_img"], out["fused_weight"] = img, w
    avg = o.psf_average([out["psf_t0"], out["psf_t1"]])
    out["psf_avg"] = avg
    out["psf_same"] = o.psf_make_same_size(avg, (25, 13, 17))
    raw_w = [o.blending_weight(DIMS, (0, 0, 0), (20, 18, 14), (0.0,) * 3, (6.0,) * 3), o.blending_weight(DIMS, (5, 3, 2), (25, 21, 17), (1.0,) * 3, (4.0,) * 3)]
    for j, r in enumerate(raw_w):
        out[f"blend{j}"] = r
    for tag, osem, smooth in (("hard", 1.0, False), ("smooth", 1.0, True), ("osem", 2.0, False)):
        for j, nw in enumerate(o.normalize_weights(raw_w, osem, smooth)):
            out[f"norm_{tag}{j}"] = nw
    np.savez_compressed(os.path.join(HERE, "prep_case.npz"), **out)
    print("wrote prep_case.npz", sum(v.nbytes for v in out.values()) // 1024, "KiB uncompressed")


if __name__ == "__main__":
    main()
