#!/usr/bin/env python3
"""Measurement of the view-materialisation row: mvd_fuse_group at the c3 grid (1024 x 1024 x 512 fused voxels) from one
rotated, anisotropically scaled raw view with fusion + deconvolution blending.  Kernel time from CUDA events inside the library
(mvd_last_fuse_group_ms), algorithmic bytes = 4 (image) + 4 (weight) written + 4 read per fused voxel; the CPU oracle is timed on a bounded
sub-box of the same grid.  python scripts/bench_materialise.py [nz ny nx]"""
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


def main():
    dims = tuple(int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (512, 1024, 1024)
    nz, ny, nx = dims
    rng = np.random.default_rng(5)
    raw_dims = (int(nz / 1.7) + 8, ny, nx)                      # anisotropic stack: z spacing 1.7
    raw = rng.random(raw_dims, dtype=np.float32) * 300
    th = np.deg2rad(7.0)
    c = np.array([nx / 2, ny / 2, nz / 2])
    R = np.array([[np.cos(th), 0, np.sin(th)], [0, 1, 0], [-np.sin(th), 0, np.cos(th)]]) @ np.diag([1, 1, 1.7])
    t = c - R @ np.array([nx / 2, ny / 2, raw_dims[0] / 2])
    fwd = np.eye(4); fwd[:3, :3] = R; fwd[:3, 3] = t
    inv = np.linalg.inv(fwd)[:3].ravel()
    bl = ((0.0, 0.0, 0.0), (12.0, 12.0, 12.0 / 1.7))
    rv = m.RawView(raw, inv, 1, bl, bl)
    psf = o.synth_psf(0, 1, (5, 5, 5), (1.0, 1.0, 1.2))
    t0 = time.perf_counter()
    dv = m.DeconViews([m.DeconView(m.FusedGroup([rv], (0, 0, 0), dims), None, psf)])
    wall = time.perf_counter() - t0
    ms = dv.last_fuse_group_ms()
    vox = nz * ny * nx
    img = dv.getImage(0)
    inside = float(np.count_nonzero(img)) / vox
    # CPU oracle on a bounded sub-box (same transform, offset bounding box)
    sub = (64, 256, 256)
    off = (nx // 2 - 128, ny // 2 - 128, nz // 2 - 32)
    t0 = time.perf_counter()
    ref_img, ref_w = o.fuse_group([raw], [inv], off, sub, 1, [bl], [bl])
    cpu_s = time.perf_counter() - t0
    got = img[off[2]:off[2] + 64, off[1]:off[1] + 256, off[0]:off[0] + 256]
    out = {"row": "view materialisation (mvd_fuse_group)", "fused_grid_zyx": dims, "raw_zyx": raw_dims, "kernel_ms": ms,
           "G_fused_voxels_per_s": vox / (ms * 1e-3) / 1e9, "algorithmic_GBs_12B_per_voxel": 12.0 * vox / (ms * 1e-3) / 1e9,
           "fraction_of_6542.7_GBs": 12.0 * vox / (ms * 1e-3) / 1e9 / 6542.7, "inside_fraction": inside,
           "wall_s_incl_context_and_h2d_of_raw": wall, "bit_exact_vs_oracle_on_sub_box": bool(np.array_equal(got, ref_img)),
           "cpu_oracle": {"G_fused_voxels_per_s": np.prod(sub) / cpu_s / 1e9, "sample_zyx": sub, "seconds": cpu_s, "cores": os.cpu_count(), "kind": "port (numpy)"}}
    print(json.dumps(out))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "materialise.json"), "w") as f:
        json.dump(out, f, indent=1)
    dv.close()


if __name__ == "__main__":
    main()
