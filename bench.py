#!/usr/bin/env python3
"""
bench.py -- headline benchmark of the B200 multi-view deconvolution path.

    python bench.py --gpus N --steps K --warmup W              (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W
    python bench.py --config c4|c5 ...                         (the 8-GPU configurations of BASELINE.json)

metric    voxel*view*iterations / s   (BASELINE.json)
workload  c3 (default): 4-view 1024x1024x512 efficient-Bayesian OSEM deconvolution, PSF 25x19x25, lambda = 0, kernel spectra
          resident in HBM; a (y x z) grid of boxes with per-view-update halo exchange for N > 1 (strong scaling)
step      ONE OSEM iteration = one view update for each view over the whole volume
value     kernel-only throughput, inputs resident in HBM, CUDA events on the context's stream, max over ranks
e2e       the same metric through the public host API with HOST buffers: upload of all views, weight masks generated on the device,
          PSF -> kernel derivation and spectra, PsiInit, `e2e_iterations` iterations, download of psi, wall clock of the whole job
parity    after the timed loop a crop of psi that straddles a rank / tile boundary is advanced by `views` further view updates on the
          GPUs and, from the same state, by the float64 CPU oracle on crop + halo; relL2 / max-abs of the crop core go into the JSON line
          and the run FAILS above the tolerance
roofline  dominant pass kernel: algorithmic bytes (SURVEY 8d per-voxel figure x useful voxels per launch) / CUDA-event time
cpu_baseline / --impl reference : the numpy/scipy oracle (CPU restatement of the reference path; the reference itself is
          Java and cannot run here) on a bounded 256x256x128 sample of the same workload, all host cores.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # dims (z, y, x); psf / sigma (x, y, z); tilt: PSF tilt step in degrees ((v - V//2) * tilt) instead of v * 180 / V
    "c1": dict(dims=(128, 256, 256), views=4, psf=(25, 19, 25), sigma=(1.5, 1.5, 4.0), lam=0.0, ptype=2, tilt=None, iters=10),
    "c2": dict(dims=(256, 512, 512), views=6, psf=(25, 19, 25), sigma=(1.5, 1.5, 4.0), lam=0.006, ptype=2, tilt=None, iters=10),
    "c3": dict(dims=(512, 1024, 1024), views=4, psf=(25, 19, 25), sigma=(1.5, 1.5, 4.0), lam=0.0, ptype=2, tilt=None, iters=10),
    "c4": dict(dims=(1024, 2048, 2048), views=8, psf=(25, 19, 25), sigma=(1.5, 1.5, 4.0), lam=0.0, ptype=3, tilt=None, iters=10),
    "c5": dict(dims=(768, 1536, 1536), views=7, psf=(31, 31, 61), sigma=(2.5, 2.5, 9.0), lam=0.0, ptype=0, tilt=4.0, iters=30),
}
WORKLOAD_TEXT = {
    "c1": "c1: 4-view 256x256x128 efficient-Bayesian OSEM, PSF 25x19x25, lambda=0",
    "c2": "c2: 6-view 512x512x256 efficient-Bayesian OSEM + Tikhonov lambda=0.006, PSF 25x19x25",
    "c3": "c3: 4-view 1024x1024x512 efficient-Bayesian OSEM, PSF 25x19x25, lambda=0, spectra resident in HBM",
    "c4": "c4: 8-view 2048x2048x1024 INDEPENDENT (classic multi-view RL), PSF 25x19x25, weight masks generated on the device",
    "c5": "c5: 7-view 1536x1536x768 OPTIMIZATION_II, anisotropic 31x31x61 PSFs tilted in 4 degree steps (large-kernel path)",
}
B_ALG = 92.0                                   # algorithmic bytes per voxel*view*iteration
PASS_BYTES = [8, 8, 12, 8, 12, 8, 12, 8, 16]   # per real voxel, passes P1..P9 (sum = 92)
PASS_NAMES = ["P1 x forward (x kernel, X_FWD; c3: x_kernel_w)", "P2 + P6 y forward (col_kernel<COL_FWD>)", "P3 + P7 z convolution (col_kernel<COL_CONV>)",
              "P4 + P8 y inverse (col_kernel<COL_INV>)", "P5 x quotient (x kernel, X_RATIO; c3: x_kernel_w)", "P6 y forward (col_kernel<COL_FWD>)",
              "P7 z convolution K2 (col_kernel<COL_CONV>)", "P8 y inverse (col_kernel<COL_INV>)", "P9 x update (x_kernel<X_UPDATE>)"]
SEED = 20263
BLEND_RANGE, BLEND_BORDER = 12.0, 0.0
PSI_SIGMA = 5.0                                # PsiInitBlurredFused default (DeconvolutionGUI.java:149)


# ------------------------------------------------------------------------------------------------------------------
# synthetic inputs, generated per box
# ------------------------------------------------------------------------------------------------------------------
def splitmix64(x):
    with np.errstate(over="ignore"):
        z = np.asarray(x, dtype=np.uint64) + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def rng_uniform(seed, stream, index):
    with np.errstate(over="ignore"):
        key = splitmix64(np.uint64(seed) * np.uint64(0x632BE59BD9B4E019) + np.uint64(stream))
        z = splitmix64(key + np.asarray(index, dtype=np.uint64))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def synth_psf(view, num_views, size_xyz, sigma_xyz, tilt_step=None):
    kx, ky, kz = size_xyz
    theta = math.radians(view * 180.0 / num_views if tilt_step is None else (view - num_views // 2) * tilt_step)
    x = np.arange(kx, dtype=np.float64) - kx // 2
    y = np.arange(ky, dtype=np.float64) - ky // 2
    z = np.arange(kz, dtype=np.float64) - kz // 2
    Z, Y, X = np.meshgrid(z, y, x, indexing="ij")
    c, s = math.cos(theta), math.sin(theta)
    xr, zr = c * X + s * Z, -s * X + c * Z
    g = np.exp(-0.5 * ((xr / sigma_xyz[0]) ** 2 + (Y / sigma_xyz[1]) ** 2 + (zr / sigma_xyz[2]) ** 2))
    return (g / g.sum()).astype(np.float32)


def truth_box(dims_zyx, y0, y1, z0, z1, seed):
    """background 100 + point beads (count N/8192, amplitude U[500,4000]) restricted to rows [y0, y1) x planes [z0, z1)."""
    nz, ny, nx = dims_zyx
    nb = max(1, nz * ny * nx // 8192)
    idx = np.arange(nb, dtype=np.uint64)
    px = np.minimum((rng_uniform(seed, 1, idx) * nx).astype(np.int64), nx - 1)
    py = np.minimum((rng_uniform(seed, 2, idx) * ny).astype(np.int64), ny - 1)
    pz = np.minimum((rng_uniform(seed, 3, idx) * nz).astype(np.int64), nz - 1)
    amp = (500.0 + 3500.0 * rng_uniform(seed, 4, idx)).astype(np.float32)
    out = np.full((z1 - z0, y1 - y0, nx), 100.0, dtype=np.float32)
    keep = (pz >= z0) & (pz < z1) & (py >= y0) & (py < y1)
    np.add.at(out, (pz[keep] - z0, py[keep] - y0, px[keep]), amp[keep])
    return out


def coverage_box(dims_zyx, view):
    nz, ny, nx = dims_zyx
    mn, mx = [0, 0, 0], [nx - 1, ny - 1, nz - 1]
    side = view % 6
    d = side // 2
    ext = (nx, ny, nz)[d]
    if side % 2 == 0:
        mn[d] = ext // 8
    else:
        mx[d] = ext - 1 - ext // 8
    return mn, mx


def make_box_images(torch, lib, W, y0, y1, z0, z1, device):
    """observed images of all views on rows [y0, y1) x planes [z0, z1) of the global volume as torch device tensors:
    img_v = max(1, truth (*) PSF_v) inside view v's coverage box, 0 outside (minValueImg / outsideValueImg)."""
    dims, V = W["dims"], W["views"]
    nz, ny, nx = dims
    ky, kz = W["psf"][1], W["psf"][2]
    m0, m1 = max(0, z0 - kz), min(nz, z1 + kz)            # margins so the box edges see true neighbours
    n0, n1 = max(0, y0 - ky), min(ny, y1 + ky)
    truth = truth_box(dims, n0, n1, m0, m1, SEED)
    psfs = [synth_psf(v, V, W["psf"], W["sigma"], W["tilt"]) for v in range(V)]
    imgs = []
    for v in range(V):
        blurred = np.ascontiguousarray(lib.convolve(truth, psfs[v], "mirror", device=device)[z0 - m0:z1 - m0, y0 - n0:y1 - n0])
        mn, mx = coverage_box(dims, v)
        t = torch.from_numpy(blurred).to(f"cuda:{device}")
        del blurred
        t.clamp_(min=1.0)                                     # minValueImg
        zs, ze = max(mn[2], z0) - z0, min(mx[2] + 1, z1) - z0
        ys, ye = max(mn[1], y0) - y0, min(mx[1] + 1, y1) - y0
        if ze > zs and ye > ys:                               # outsideValueImg = 0 outside the coverage box
            t[:zs] = 0; t[ze:] = 0
            t[:, :ys] = 0; t[:, ye:] = 0
            t[:, :, :mn[0]] = 0; t[:, :, mx[0] + 1:] = 0
        else:
            t.zero_()
        imgs.append(t)
    del truth
    torch.cuda.synchronize()
    return psfs, imgs


# ------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (in-process NVML, 10 ms period; nvidia-smi fallback)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, device):
        self.device = device
        self.sm, self.mask, self.max_mhz = [], 0, None
        self.stop_flag = threading.Event()
        self.thread = None
        self.nvml = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.device]) if vis and vis.split(",")[0].isdigit() else self.device
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nvml = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
        except Exception:
            self.nvml = None

    def _loop(self):
        n = self.nvml
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
                try:
                    self.mask |= int(n.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    self.mask |= int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:
                pass
            time.sleep(0.01)

    def stop(self):
        if self.nvml is None:
            try:
                q = "clocks.sm,clocks.max.sm"
                out = subprocess.run(["nvidia-smi", f"--id={self.device}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=10).stdout.strip().split(",")
                return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "reasons": ["nvml unavailable: single nvidia-smi sample after the region"], "samples": 1}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"], "samples": 0}
        self.stop_flag.set()
        self.thread.join(timeout=1.0)
        reasons = sorted(name for bit, name in self.REASONS.items() if self.mask & bit)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(self.sm)}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(pass_index):
    """per-launch dram bytes of the pass kernel from the committed ncu capture (profiles/ncu_traffic.json), or None"""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f).get(f"P{pass_index + 1}")
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------------------------
# CPU leg: the oracle (numpy/scipy restatement of the reference CPU path) on a bounded sample
# ------------------------------------------------------------------------------------------------------------------
def load_oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import mvdecon_oracle as o
    return o


def cpu_sample_setup(name):
    o = load_oracle()
    W = WORKLOADS[name]
    dims, V = W["dims"], W["views"]
    sdims = (min(dims[0], 128), min(dims[1], 256), min(dims[2], 256))
    ds = o.make_synthetic(sdims, V, seed=SEED, psf_size_xyz=W["psf"], psf_sigma_xyz=W["sigma"], tilt_step_deg=W["tilt"])
    k1, k2 = o.derive_kernels(ds.psfs, W["ptype"])
    fused, mx, avg = o.psi_init_fused_stats(ds.images, ds.weights)
    views = [o.OracleView(ds.images[v], ds.weights[v], k1[v], k2[v], float(mx[v])) for v in range(V)]
    psi0 = np.where(fused > 0, fused, np.float32(avg)).astype(np.float32)
    return o, views, psi0, W["lam"], sdims


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if not args.child:
        # torchrun exports OMP_NUM_THREADS=1 (and friends) to its workers; the CPU arm must see the box's host cores exactly like a
        # plain `python bench.py --impl reference` does, so it runs in a child with the launcher's variables removed and the full
        # affinity mask restored
        drop = ("OMP_", "MKL_", "OPENBLAS_", "NUMEXPR_", "TORCHELASTIC_", "TORCH_NCCL", "NCCL_", "MASTER_", "GROUP_", "ROLE_", "LOCAL_")
        env = {k: v for k, v in os.environ.items() if not k.startswith(drop) and k not in ("RANK", "WORLD_SIZE")}

        def full_affinity():
            try:
                os.sched_setaffinity(0, range(os.cpu_count() or 1))
            except Exception:
                pass

        cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--child", "--gpus", str(args.gpus), "--steps", str(args.steps),
               "--warmup", str(args.warmup), "--config", args.config]
        sys.exit(subprocess.run(cmd, env=env, preexec_fn=full_affinity).returncode)
    name = args.config
    V = WORKLOADS[name]["views"]
    o, views, psi0, lam, sdims = cpu_sample_setup(name)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    psi = psi0
    for _ in range(args.warmup):
        psi, _st = o.run_iterations_seq(psi, views, 1, lam, dtype=np.float32)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        psi, _st = o.run_iterations_seq(psi, views, 1, lam, dtype=np.float32)
    dt = time.perf_counter() - t0
    vox = sdims[0] * sdims[1] * sdims[2]
    val = vox * V * args.steps / dt
    sample = f"{sdims[2]}x{sdims[1]}x{sdims[0]} sub-volume of the workload, {V} views, float32 oracle (numpy + scipy.fft, workers={cores}), 1 iteration per step"
    print(json.dumps({
        "impl": "reference", "metric": "voxel*view*iterations/s", "value": val, "unit": "voxel*view*iterations/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_TEXT[name], "step": "one OSEM iteration (all views) on the CPU sample"},
        "cpu_baseline": {"value": val, "unit": "voxel*view*iterations/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "voxel*view*iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference is Java (no JVM in this image): this arm times the CPU restatement in oracle/ (parity unpinned, see DESIGN.md); "
                "it runs in a child process without the launcher's OMP_NUM_THREADS so that N > 1 launches see the same host cores as N = 1",
    }), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# in-bench parity: a crop that straddles a rank / tile boundary against the float64 oracle
# ------------------------------------------------------------------------------------------------------------------
def parity_region(W, nviews, yc, zc, core=32):
    """(region, core) as [(lo, hi)] * 3 in (z, y, x) order: the core cube centred at (zc, yc, nx/2), the region = core + nviews * (k - 1)
    per side, clipped to the volume (a clipped face is a true volume face, where the oracle's own boundary handling is exact)."""
    nz, ny, nx = W["dims"]
    kx, ky, kz = W["psf"]
    cen = (zc, yc, nx // 2)
    halo = (nviews * (kz - 1), nviews * (ky - 1), nviews * (kx - 1))
    cores, regs = [], []
    for c, h, n in zip(cen, halo, (nz, ny, nx)):
        lo = min(max(0, c - core // 2), n - core)
        cores.append((lo, lo + core))
        regs.append((max(0, lo - h), min(n, lo + core + h)))
    return regs, cores


def gather_region(torch, dist, local_t, own, loc0, region):
    """region-shaped float32 tensor holding this job's values: every rank fills the part of the region its own box covers from its
    local array (local_t[z - loc0[0], y - loc0[1], x]) and the parts are summed over the ranks (the other ranks contribute exact zeros)."""
    (rz0, rz1), (ry0, ry1), (rx0, rx1) = region
    out = torch.zeros((rz1 - rz0, ry1 - ry0, rx1 - rx0), dtype=torch.float32, device=local_t.device)
    (zlo, zhi), (ylo, yhi) = own
    a0, a1, b0, b1 = max(zlo, rz0), min(zhi, rz1), max(ylo, ry0), min(yhi, ry1)
    if a1 > a0 and b1 > b0:
        out[a0 - rz0:a1 - rz0, b0 - ry0:b1 - ry0, :] = local_t[a0 - loc0[0]:a1 - loc0[0], b0 - loc0[1]:b1 - loc0[1], rx0:rx1]
    if dist is not None:
        dist.all_reduce(out)
    return out


def oracle_region_update(W, psfs, region, psi_reg, img_regs, maxv, nviews):
    """`nviews` view updates of the float64 oracle on the region (kernels derived by the oracle itself from the raw PSFs with the
    reference's normalisation, weights = the oracle's blending + NormalizingRandomAccess restatement on the region's coordinates)."""
    o = load_oracle()
    V = W["views"]
    k1, k2 = o.derive_kernels(psfs, W["ptype"])
    rdims = tuple(hi - lo for lo, hi in region)
    off = (region[2][0], region[1][0], region[0][0])                 # (x, y, z) offset of the region
    raw = []
    for v in range(V):
        mn, mx = coverage_box(W["dims"], v)
        raw.append(o.blending_weight(rdims, [mn[d] - off[d] for d in range(3)], [mx[d] - off[d] for d in range(3)],
                                     (BLEND_BORDER,) * 3, (BLEND_RANGE,) * 3))
    ws = o.normalize_weights(raw, 1.0, False)
    psi = psi_reg
    for v in range(nviews):
        view = o.OracleView(img_regs[v], ws[v], k1[v], k2[v], float(maxv[v]))
        psi, _, _ = o.view_update_whole(psi, view, W["lam"], dtype=np.float64)
    return psi


# ------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import mvrecon_b200 as m

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    lib = m.lib()
    name = args.config
    W = WORKLOADS[name]
    dims, V, psf_xyz, lam, ptype = W["dims"], W["views"], W["psf"], W["lam"], W["ptype"]
    nz, ny, nx = dims
    from mvrecon_b200 import sharding
    # exchange scheme 1 (psi by k1/2, then the quotient by k2/2; interior halo max(k1/2, k2/2)); BENCH_SCHEME=0 selects the single-exchange scheme
    scheme = int(os.environ.get("BENCH_SCHEME", "1")) if world > 1 else 0
    Hy, Hz = ((psf_xyz[1] - 1) // 2, (psf_xyz[2] - 1) // 2) if scheme == 1 else (psf_xyz[1] - 1, psf_xyz[2] - 1)
    py, pz = sharding.grid_for(world, ny, nz, (psf_xyz[1] - 1) // 2, (psf_xyz[2] - 1) // 2, lib.supported_fft_lengths(), scheme)
    if os.environ.get("BENCH_GRID"):                         # experiments: "PYxPZ"
        py, pz = (int(x) for x in os.environ["BENCH_GRID"].split("x"))
    ry, rz = rank // pz, rank % pz
    ylo, yhi = sharding.slab_range(ny, py, ry)
    lo, hi = sharding.slab_range(nz, pz, rz)
    y0, y1 = sharding.extended_range(ylo, yhi, ny, Hy)
    z0, z1 = sharding.extended_range(lo, hi, nz, Hz)
    local_shape = (z1 - z0, y1 - y0, nx)

    psfs, imgs = make_box_images(torch, lib, W, y0, y1, z0, z1, local)
    shard = None if pz == 1 else (lo, hi, z0, z1 - z0)
    shard_y = None if py == 1 else (ylo, yhi, y0, y1 - y0)

    comm = None
    if world > 1:                                            # one NCCL communicator per process, reused by every context of this run
        ids = [lib.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        comm = lib.comm_create(ids[0], world, rank, local)

    def build(images, async_upload=False):
        """DeconViews from images only: the weight masks are generated on the device (BlendingRealRandomAccess of every view's coverage
        box + NormalizingRandomAccess), the halo exchange is the library's own (peer stores / NCCL on the compute stream)"""
        views = [m.DeconView(im, None, psfs[v], m.PSFTYPE(ptype)) for v, im in enumerate(images)]
        dv_ = m.DeconViews(views, device=local, lambda_=lam, shard=shard, shard_y=shard_y, global_dims_zyx=dims, async_upload=async_upload,
                           exchange_scheme=scheme)
        for v in range(V):
            mn, mx = coverage_box(dims, v)
            dv_.makeBlendingWeights(v, mn, mx, (BLEND_BORDER,) * 3, (BLEND_RANGE,) * 3)
        dv_.normalizeWeights(1.0, False)
        if comm is not None:
            dv_.comm_attach(comm, py, pz)
        return dv_

    # ---------------- kernel-only leg: everything resident --------------------------------------------------------
    dv = build([m.DeviceArray.from_torch(im) for im in imgs])
    info = dv.tile_info()
    init = m.PsiInitBlurredFused(PSI_SIGMA)                  # on the device; on a sharded context avg / max[] are all-reduced by the library
    dec = m.MultiViewDeconvolutionSeq(dv, 0, init)
    if not dec.initWasSuccessful():
        raise RuntimeError("PsiInit failed")
    maxv = [float(x) for x in init.getMax()]
    stream = torch.cuda.ExternalStream(dv.stream_handle(), device=f"cuda:{local}")
    transport = dv.exchange_transport() if world > 1 else "none"

    def one_iteration(nviews=V):
        for v in range(nviews):
            dv.enqueue_view_update(v)            # with a communicator attached the library exchanges the halos itself

    for _ in range(args.warmup):
        one_iteration()
    dv.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if os.environ.get("BENCH_EXCHANGE_ONLY") and world > 1:          # development: cost of the bare psi halo exchange
        # BENCH_EXCHANGE_GAP_US: busy-wait of that length on the compute stream before every exchange (what a pass kernel does to the
        # links: they sit idle in between) -- exchange time is then measured per exchange, gaps excluded
        gap_us = float(os.environ.get("BENCH_EXCHANGE_GAP_US", "0"))
        cycles = int(gap_us * 1.9e3)
        n_ex = 50
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_ex)]
        for _ in range(5):
            dv.exchange_halos()
        dv.synchronize(); dist.barrier()
        for a, b in evs:
            if cycles:
                with torch.cuda.stream(stream):
                    torch.cuda._sleep(cycles)
            a.record(stream)
            dv.exchange_halos()
            b.record(stream)
        evs[-1][1].synchronize()
        per = sorted(a.elapsed_time(b) * 1e3 for a, b in evs)
        t = torch.tensor([sum(per) / n_ex, per[n_ex // 2], per[0], per[-1]], device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(json.dumps({"exchange_only_us": float(t[0].item()), "median_us": float(t[1].item()), "min_us": float(t[2].item()),
                              "max_us": float(t[3].item()), "gap_us": gap_us, "transport": transport, "grid": [py, pz], "halo": [Hy, Hz],
                              "local_box_zyx": list(local_shape)}))
        dv.close(); dist.barrier(); dist.destroy_process_group()
        return
    dv.set_profiling(True)
    dv.pass_times(reset=True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                 # NVML initialisation takes 10 - 100 ms on this rank only ...
    if world > 1:
        dist.barrier()                  # ... so the ranks meet again AFTER it: the timed region starts on all ranks together
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        one_iteration()
    dv.synchronize()
    torch.cuda.synchronize()
    e1.record(stream)
    e1.synchronize()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    aux_ms, aux_n = dv.aux_times(reset=False)
    pass_ms, pass_n = dv.pass_times(reset=True)
    dv.aux_times(reset=True)
    dv.set_profiling(False)
    nvu = max(V * args.steps, 1)
    aux = {"quotient_exchange_ms_per_view_update": round(aux_ms[0] / nvu, 4), "between_view_updates_ms_per_view_update": round(aux_ms[1] / nvu, 4),
           "joins_and_clears_ms_per_view_update": round(aux_ms[2] / nvu, 4)}
    rank_pass_ms = None
    if world > 1:
        t = torch.tensor([ms], device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        # time every rank spent inside its nine passes per step (CUDA events around the launches): the rest of the step is exchange + waiting
        # for the slowest neighbour
        mine = torch.zeros(world, device=f"cuda:{local}")
        mine[rank] = sum(pass_ms) / max(args.steps, 1)
        dist.all_reduce(mine)
        rank_pass_ms = [round(float(x), 3) for x in mine.tolist()]
    vox = nz * ny * nx
    value = vox * V * args.steps / (ms * 1e-3)

    # ---------------- parity: GPU vs float64 oracle on a crop that straddles a rank / tile boundary ----------------
    parity = None
    if not args.skip_parity:
        big_k = max(psf_xyz) > 25
        pv = min(V, 2 if big_k else 4)                       # view updates checked (bounds the halo the CPU has to carry)
        yc = sharding.slab_range(ny, py, py // 2)[0] if py > 1 else ny // 2        # N = 1: the y tiles of the plan meet near ny / 2
        zc = sharding.slab_range(nz, pz, pz // 2)[0] if pz > 1 else nz // 2
        region, core = parity_region(W, pv, yc, zc)
        own = ((lo, hi), (ylo, yhi))

        def psi_tensor():
            return torch.as_tensor(m.RawDeviceBuffer(dv.psi_device_ptr(), local_shape), device=f"cuda:{local}")

        with torch.cuda.stream(stream):
            before = gather_region(torch, dist, psi_tensor(), own, (z0, y0), region)
            img_regs = [gather_region(torch, dist, imgs[v], own, (z0, y0), region) for v in range(pv)]
        stream.synchronize()
        one_iteration(pv)
        dv.synchronize()
        with torch.cuda.stream(stream):
            after = gather_region(torch, dist, psi_tensor(), own, (z0, y0), region)
        stream.synchronize()
        finite_local = bool(torch.isfinite(psi_tensor()).all().item())
        if rank == 0:
            tp = time.perf_counter()
            ref = oracle_region_update(W, psfs, region, before.cpu().numpy(), [t.cpu().numpy() for t in img_regs], maxv, pv)
            sl = tuple(slice(c[0] - r[0], c[1] - r[0]) for c, r in zip(core, region))
            got = after.cpu().numpy()[sl].astype(np.float64)
            want = ref[sl].astype(np.float64)
            rel = float(np.linalg.norm(got - want) / max(np.linalg.norm(want), 1e-300))
            mabs = float(np.abs(got - want).max())
            tol_rel, tol_abs = 2e-6, 1e-3 * float(np.abs(want).max())
            parity = {"relL2": rel, "max_abs": mabs, "tol_relL2": tol_rel, "tol_max_abs": tol_abs, "ok": bool(rel <= tol_rel and mabs <= tol_abs),
                      "view_updates": pv, "core_zyx": [list(c) for c in core], "region_zyx": [list(r) for r in region],
                      "straddles": ("rank boundary y=%d z=%d" % (yc, zc)) if world > 1 else ("tile boundary near y=%d" % yc),
                      "oracle": "float64 CPU oracle from the same float32 state (kernels and weight masks re-derived by the oracle)",
                      "seconds": round(time.perf_counter() - tp, 2)}
        del before, after, img_regs
    else:
        finite_local = bool(np.isfinite(dec.getPSI()).all())
    finite = finite_local
    if world > 1:
        ft = torch.tensor([1.0 if finite_local else 0.0], device=f"cuda:{local}")
        dist.all_reduce(ft, op=dist.ReduceOp.MIN)
        finite = bool(ft.item() > 0.5)
    dv.close()
    del dec

    # ---------------- e2e leg: public host API, host buffers, copies inside the timed region --------------------
    e2e_iters = args.e2e_iterations
    if args.skip_e2e:
        if rank == 0:
            print(json.dumps({"metric": "voxel*view*iterations/s", "value": value, "n_gpus": world, "ms_per_step": ms / args.steps,
                              "note": "profiling run (--skip-e2e): not a bench line", "transport": transport, "parity": parity,
                              "fft_tile_xyz": info["tile_dims_xyz"], "tiles_per_gpu": info["num_tiles"], "pass_ms_per_step_by_rank": rank_pass_ms, "aux_rank0": aux,
                              "all_passes_ms_per_launch": [round(a / max(b, 1), 4) for a, b in zip(pass_ms, pass_n)]}))
        if comm is not None:
            comm.close()
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
        if parity is not None and not parity["ok"]:
            sys.exit(1)
        return
    host = []
    for im in imgs:
        hi_ = torch.empty(im.shape, dtype=torch.float32, pin_memory=True)
        hi_.copy_(im)
        host.append(hi_.numpy())
    out_pinned = torch.empty(local_shape, dtype=torch.float32, pin_memory=True)
    del imgs
    torch.cuda.empty_cache()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    dv = build(host, async_upload=True)                      # H2D of all views (copy stream), weights on the device, PSF -> kernels, spectra, exchange
    stream = torch.cuda.ExternalStream(dv.stream_handle(), device=f"cuda:{local}")
    dec = m.MultiViewDeconvolutionSeq(dv, 0, m.PsiInitBlurredFused(PSI_SIGMA))
    t_setup = time.perf_counter() - t0
    for _ in range(e2e_iters):
        one_iteration()
    t_enq = time.perf_counter() - t0
    out = dec.getPSI(out=out_pinned.numpy())                 # D2H into page-locked host memory
    t_e2e = time.perf_counter() - t0
    t_marks = [round(t_setup, 4), round(t_enq, 4), round(t_e2e, 4)]
    if world > 1:
        t = torch.tensor([t_e2e], device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())
    e2e_val = vox * V * e2e_iters / t_e2e
    local_vox = local_shape[0] * local_shape[1] * local_shape[2]
    h2d = V * local_vox * 4 * world / e2e_iters
    d2h = local_vox * 4 * world / e2e_iters
    finite = finite and bool(np.isfinite(out).all())
    launches = info["launches_per_view_update"] * V * args.steps
    dv.close()

    if rank == 0:
        peak, peak_src = measured_peak()
        # dominant KERNEL = the kernel name with the largest accumulated time over the step; P2 / P6, P3 / P7 and P4 / P8 are the same
        # kernel launched twice per view update (the ncu launch list under profiles/ groups them the same way)
        groups = [(0,), (1, 5), (2, 6), (3, 7), (4,), (8,)]
        gi = int(np.argmax([sum(pass_ms[i] for i in g) for g in groups]))
        dom = groups[gi][0]
        dom_ms = sum(pass_ms[i] for i in groups[gi])
        per_launch_ms = dom_ms / max(sum(pass_n[i] for i in groups[gi]), 1)
        useful_vox_per_launch = (hi - lo) * (yhi - ylo) * nx / info["num_tiles"]
        achieved = PASS_BYTES[dom] * useful_vox_per_launch / (per_launch_ms * 1e-3) / 1e9
        per_pass_frac = [round(PASS_BYTES[i] * useful_vox_per_launch / (pass_ms[i] / max(pass_n[i], 1) * 1e-3) / 1e9 / peak, 4) if pass_ms[i] > 0 else None
                         for i in range(9)]
        cpu = None
        if world == 1 and not args.skip_cpu:
            o, views, cpsi, clam, sdims = cpu_sample_setup(name)
            cores = os.cpu_count() or 1
            cpsi, _ = o.run_iterations_seq(cpsi, views, 1, clam, dtype=np.float32)      # warm-up (plans, page faults)
            tc = time.perf_counter()
            n_it = 2
            o.run_iterations_seq(cpsi, views, n_it, clam, dtype=np.float32)
            tc = time.perf_counter() - tc
            cpu = {"value": sdims[0] * sdims[1] * sdims[2] * V * n_it / tc, "unit": "voxel*view*iterations/s", "cores": cores, "kind": "port",
                   "sample": f"{sdims[2]}x{sdims[1]}x{sdims[0]} sub-volume, {V} views, {n_it} iterations, float32 oracle (numpy + scipy.fft, workers={cores})"}
        line = {
            "metric": "voxel*view*iterations/s", "value": value, "unit": "voxel*view*iterations/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD_TEXT[name], "step": f"one OSEM iteration = {V} view updates over the whole volume",
                       "fft_tile_xyz": info["tile_dims_xyz"], "tiles_per_gpu": info["num_tiles"], "fft_box_over_useful_voxels": round(info["fft_volume_ratio"], 4),
                       "sharding": "none" if world == 1 else f"{py} x {pz} (y x z) boxes, exchange scheme {scheme} (" + ("psi by k1/2 before + quotient spectrum by k2/2 inside" if scheme == 1 else "psi by k1/2 + k2/2 after") + f" every view update; local halo {Hy} rows / {Hz} planes), enqueued on the compute stream by the library, transport {transport}",
                       "inputs": "images resident; weight masks (cosine blending + NormalizingRandomAccess) and psi0 (PsiInitBlurredFused, sigma 5) generated on the device; "
                                 "PSF normalisation = the reference's (AdjustInput.sumImg double count, T = Threads.numThreads() of this host)",
                       "l2_flush": "not needed: every pass streams >= 1.2 GB (inputs larger than the 126 MB L2)",
                       "roofline_fraction_92B": value * B_ALG / (peak * 1e9 * world), "output_finite": finite},
            "parity_relL2": None if parity is None else parity["relL2"], "parity": parity,
            "roofline": {"bound": "hbm", "kernel": PASS_NAMES[dom], "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(dom) if list(info["tile_dims_xyz"]) == [1080, 540, 540] else None, "peak_source": peak_src, "algorithmic_bytes_per_voxel": PASS_BYTES[dom],
                         "ms_per_launch": per_launch_ms, "share_of_step": dom_ms / max(sum(pass_ms), 1e-9),
                         "launches_per_view_update": len(groups[gi]), "frac_by_pass_P1_P9": per_pass_frac,
                         "all_passes_ms_per_launch": [round(a / max(b, 1), 4) for a, b in zip(pass_ms, pass_n)]},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_val, "unit": "voxel*view*iterations/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "iterations": e2e_iters, "seconds": t_e2e, "host_marks_s_rank0": t_marks,
                    "what": "DeconViews(page-locked host images, async upload on a copy stream) + device weight masks + PSF->kernel derivation + spectra + PsiInit + iterations + getPSI(), wall clock; bytes amortised per iteration"},
            "gpu_launches": launches, "clocks": clocks,
        }
        if rank_pass_ms is not None:
            line["pass_ms_per_step_by_rank"] = rank_pass_ms      # the step minus this = halo exchange + waiting for the slowest neighbour
            line["aux_rank0"] = aux
        print(json.dumps(line), flush=True)
    if comm is not None:
        comm.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0 and (not finite or (parity is not None and not parity["ok"])):
        sys.exit(1)                                          # a fast but wrong (or non-finite) result is not a bench value


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--e2e-iterations", type=int, default=10)
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs: kernel-only leg only")
    ap.add_argument("--skip-cpu", action="store_true", help="profiling runs: no CPU baseline leg")
    ap.add_argument("--skip-parity", action="store_true", help="profiling runs: no oracle check of the result")
    ap.add_argument("--child", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
