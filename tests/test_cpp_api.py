"""The C++ host mirror of the reference's class interface (include/mvdecon.hpp) driven by a C++ program, linked against the CPU emulator
build of the library: kernels, psi after two iterations, per-view statistics and the block operator must equal the oracle."""
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_mirror_runs_the_loop(hostemu_lib, oracle, tmp_path):
    exe = tmp_path / "cpp_api_test"
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "host", "cpp_api_test.cpp"),
                        "-o", str(exe), hostemu_lib.path, "-Wl,-rpath," + os.path.dirname(hostemu_lib.path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    dims, V, iters, lam = (20, 24, 28), 3, 2, 0.006
    ds = oracle.make_synthetic(dims, V, seed=17, psf_size_xyz=(5, 3, 5), psf_sigma_xyz=(1.0, 0.8, 1.4), bead_density=256)
    views, psi0, avg = oracle.make_oracle_views(ds, oracle.EFFICIENT_BAYESIAN)
    with open(tmp_path / "meta.txt", "w") as f:
        f.write(f"{dims[2]} {dims[1]} {dims[0]} {V} 5 3 5 {oracle.EFFICIENT_BAYESIAN} {lam} {iters}\n")
    for v in range(V):
        ds.images[v].astype(np.float32).tofile(tmp_path / f"img{v}.f32")
        ds.weights[v].astype(np.float32).tofile(tmp_path / f"w{v}.f32")
        ds.psfs[v].astype(np.float32).tofile(tmp_path / f"psf{v}.f32")
    psi0.astype(np.float32).tofile(tmp_path / "psi0.f32")
    np.array([v.max_intensity for v in views], np.float32).tofile(tmp_path / "max.f32")
    r = subprocess.run([str(exe), str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stderr
    for v in range(V):
        k1 = np.fromfile(tmp_path / f"k1_{v}.f32", np.float32).reshape(views[v].kernel1.shape)
        k2 = np.fromfile(tmp_path / f"k2_{v}.f32", np.float32).reshape(views[v].kernel2.shape)
        assert oracle.rel_l2(k1, views[v].kernel1) < 1e-6 and oracle.rel_l2(k2, views[v].kernel2) < 1e-5
    ref, stats = oracle.run_iterations_seq(psi0, views, iters, lam, dtype=np.float64)
    got = np.fromfile(tmp_path / "psi_out.f32", np.float32).reshape(dims)
    assert oracle.rel_l2(got, ref) < 4e-6
    lines = [tuple(float(x) for x in l.split()) for l in open(tmp_path / "stats.txt").read().splitlines()]
    assert len(lines) == iters * V + 1 + V
    for (s, m), (_, _, ws, wm) in zip(lines[:iters * V], stats):
        assert abs(s - ws) <= 1e-4 * max(1.0, abs(ws)) + 1e-2 and abs(m - wm) <= 1e-4 * max(1.0, abs(wm))
    # block operator on the whole volume as one block = one whole-volume update of view 0 (mirror / constant-1 extension)
    one, s0, m0 = oracle.view_update_whole(psi0, views[0], lam, dtype=np.float64)
    blk = np.fromfile(tmp_path / "block_out.f32", np.float32).reshape(dims)
    assert oracle.rel_l2(blk, one) < 2e-6
    # the block-wise driver over the operator (16^3 halo'd blocks, delayed paste-back) equals the whole-volume iteration
    one_it, _ = oracle.run_iterations_seq(psi0, views, 1, lam, dtype=np.float64)
    blocked = np.fromfile(tmp_path / "psi_blocked.f32", np.float32).reshape(dims)
    assert oracle.rel_l2(blocked, one_it) < 4e-6
