"""
CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  **Parity unpinned** (see below).

A numpy/scipy restatement of the reference's multi-view Richardson-Lucy /
efficient-Bayesian deconvolution hot path
(net.preibisch.mvrecon.process.deconvolution, PreibischLab/multiview-reconstruction).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module, and only as the checker.  The
product path (``multiview-reconstruction_b200``) never imports it.

PARITY UNPINNED: the reference is Java (no ``java``/``javac``/``mvn`` in this image, no
jars, no network), its FFT / Gauss back-ends are third-party artifacts that are absent
from ``/root/reference`` (``net.imglib2:imglib2-algorithm-fft`` -> ``edu.mines.jtk``,
``net.imglib2:imglib2-algorithm`` Gauss3; versions managed by parent BOM
``pom-scijava 43.0.0``, ``/root/reference/pom.xml:5-10,210-217``) and the reference ships
no golden vectors, JUnit tests or fixtures for this path.  The oracle
is therefore pinned only by (i) line-by-line restatement of the in-tree arithmetic cited
below, (ii) known-answer tests derived from those lines (tests/test_oracle.py) and
(iii) self-consistency (blocked == whole-volume, float32 vs float64 FFT).

Path shorthand in citations:
    M/ = /root/reference/src/main/java/net/preibisch/mvrecon/
    U/ = /root/reference/src/main/java/util/

All volumes are numpy arrays indexed [z, y, x] (x fastest in memory), which is the
reference's ArrayImg order (M/process/cuda/Block.java:299-308).  Where the reference
indexes dimensions (d = 0 is x), functions here take/return (x, y, z) tuples and say so.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np
import scipy.fft as sfft

# --------------------------------------------------------------------------------------
# constants  (M/process/deconvolution/MultiViewDeconvolution.java:48-60)
# --------------------------------------------------------------------------------------
OUTSIDE_VALUE_IMG = np.float32(0.0)
MIN_VALUE_IMG = np.float32(1.0)
MIN_VALUE = np.float32(0.0001)
DEFAULT_BLENDING_RANGE = 12
DEFAULT_BLENDING_BORDER = -8
MAX_DIFF_RANGE = np.float32(0.1)
SCALING_RANGE = np.float32(0.05)

# DeconViewPSF.PSFTYPE ordinal order (M/process/deconvolution/DeconViewPSF.java:52)
OPTIMIZATION_II, OPTIMIZATION_I, EFFICIENT_BAYESIAN, INDEPENDENT = range(4)
PSFTYPE_NAMES = ["OPTIMIZATION_II", "OPTIMIZATION_I", "EFFICIENT_BAYESIAN", "INDEPENDENT"]

_WORKERS = os.cpu_count() or 1


def num_threads(ij_threads: Optional[int] = None) -> int:
    """Threads.numThreads() = max(4, Prefs.getThreads())  (M/Threads.java:40).
    ImageJ's Prefs.getThreads() defaults to the number of processors."""
    if ij_threads is None:
        ij_threads = os.cpu_count() or 1
    return max(4, ij_threads)


# --------------------------------------------------------------------------------------
# FusionTools.divideIntoPortions   (M/process/fusion/FusionTools.java:1287-1329)
# --------------------------------------------------------------------------------------
def divide_into_portions(image_size: int, threads: Optional[int] = None) -> List[Tuple[int, int]]:
    """Returns [(start, loop_size)], last portion takes the remainder."""
    T = num_threads(threads)
    if image_size <= T:
        n = int(image_size)
    else:
        n = max(T, int(image_size // (64 * 64 * 64)))
    if image_size == 0:
        return []
    chunk = image_size // n
    while chunk == 0:
        n -= 1
        chunk = image_size // n
    mod = image_size % n
    out = []
    for p in range(n):
        start = p * chunk
        loop = chunk + mod if p == n - 1 else chunk
        out.append((start, loop))
    return out


# --------------------------------------------------------------------------------------
# AdjustInput.sumImg / normToSum1   (M/process/deconvolution/normalization/AdjustInput.java:52-122)
# --------------------------------------------------------------------------------------
REFERENCE_THREADS = "reference"      # quirk_threads value: the run-time default of the reference, Threads.numThreads() of this host


def sum_img(img: np.ndarray, quirk_threads=REFERENCE_THREADS) -> float:
    """Sum of all pixels in float64 (RealSum) the way AdjustInput.sumImg computes it.

    quirk_threads=REFERENCE_THREADS (default) or T
                        -> AdjustInput.java:115-119 as written: ``sum.add(sums[0])`` followed by a loop over
                           *all* sums including index 0, i.e. portion 0 (first floor(size/numPortions) pixels in
                           x-fastest order; sums[] is filled in task start order, portion 0 being the first task
                           submitted) is counted twice, for Threads.numThreads() == max(4, T)
                           (T = the host's processor count for REFERENCE_THREADS).  SURVEY 8a-6 'Quirk A'.
    quirk_threads=None  -> exact sum (what the author intended; opt-in, not the reference's behaviour).
    """
    flat = np.asarray(img, dtype=np.float64).ravel()  # C order of [z,y,x] == x fastest
    total = math.fsum(flat.tolist()) if flat.size <= 1 << 20 else float(flat.sum(dtype=np.float64))
    if quirk_threads is not None:
        if quirk_threads == REFERENCE_THREADS:
            quirk_threads = None                       # num_threads(None) = max(4, processors)
        start, loop = divide_into_portions(flat.size, quirk_threads)[0]
        total += float(flat[start:start + loop].sum(dtype=np.float64))
    return total


def norm_to_sum1(img: np.ndarray, quirk_threads=REFERENCE_THREADS) -> np.ndarray:
    """t = (float)((double)t / sum)   (AdjustInput.java:52-58). Returns a new float32 array."""
    s = sum_img(img, quirk_threads)
    return (np.asarray(img, dtype=np.float32).astype(np.float64) / s).astype(np.float32)


# --------------------------------------------------------------------------------------
# Mirror.mirror / computeInvertedKernel
#   (M/process/deconvolution/util/Mirror.java:55-132, DeconViewPSF.java:266-274)
# --------------------------------------------------------------------------------------
def mirror_axis(img: np.ndarray, axis: int) -> np.ndarray:
    """In-place swap loop of Mirror.java:96-108 restated: every position p <= dim/2 is swapped
    with dim-1-p.  For odd sizes that is a true flip.  For even size 2m the positions m-1 and m
    are swapped twice (Quirk C) so the two middle samples stay where they were."""
    n = img.shape[axis]
    out = np.flip(img, axis=axis).copy()
    if n % 2 == 0 and n >= 2:
        m = n // 2
        sl_a = [slice(None)] * img.ndim
        sl_b = [slice(None)] * img.ndim
        sl_a[axis] = m - 1
        sl_b[axis] = m
        # double swap == identity for the two middle samples
        out[tuple(sl_a)] = img[tuple(sl_a)]
        out[tuple(sl_b)] = img[tuple(sl_b)]
    return out


def compute_inverted_kernel(kernel: np.ndarray) -> np.ndarray:
    out = np.asarray(kernel, dtype=np.float32)
    for axis in range(out.ndim):
        out = mirror_axis(out, axis)
    return np.ascontiguousarray(out)


def compute_exponential_kernel(kernel: np.ndarray, num_views: int) -> np.ndarray:
    """pow by repeated float multiply (DeconViewPSF.java:256-264,276-284)."""
    k = np.asarray(kernel, dtype=np.float32)
    res = k.copy()
    for _ in range(1, num_views):
        res = (res * k).astype(np.float32)
    return res


# --------------------------------------------------------------------------------------
# U/FFTConvolution.convolve   (U/FFTConvolution.java:490-603,659-666)
# --------------------------------------------------------------------------------------
def _pad_volume(img: np.ndarray, pads, ext: str, const: float) -> np.ndarray:
    if ext == "mirror":          # Views.extendMirrorSingle == numpy 'reflect'
        # numpy reflect handles pad > n-1 by repeated reflection, like the periodic mirror strategy; a size-1 axis mirrors onto itself
        out = img
        for ax, pw in enumerate(pads):
            if pw[0] == 0 and pw[1] == 0:
                continue
            one = [(0, 0)] * img.ndim
            one[ax] = tuple(pw)
            out = np.pad(out, one, mode="reflect" if out.shape[ax] > 1 else "edge")
        return out
    if ext == "zero":
        return np.pad(img, pads, mode="constant", constant_values=0)
    if ext == "const":
        return np.pad(img, pads, mode="constant", constant_values=const)
    if ext == "periodic":
        return np.pad(img, pads, mode="wrap")
    raise ValueError(ext)


def fft_convolve(img: np.ndarray, kernel: np.ndarray, ext: str = "mirror", const: float = 1.0,
                 dtype=np.float32, workers: int = _WORKERS) -> np.ndarray:
    """y(x) = sum_t img_ext(x - t) * K(t + c), c = floor(k/2) per axis, output = size of img.

    Follows U/FFTConvolution.java:508-544 (pad to >= img + k - 1, image extended by its
    out-of-bounds strategy, kernel centre dim/2 moved to the origin with periodic wrap) and
    :584-603 (multiply spectra, inverse, un-pad).  The padded size only has to be >= img+k-1
    (any such size yields the same linear convolution), so scipy's next_fast_len replaces
    Mines-JTK's nfftFast.  dtype selects float32 (emulates the reference's float FFT) or
    float64 (truth) arithmetic.  complexConjugate == false (U/FFTConvolution.java:94).
    """
    img = np.asarray(img)
    kernel = np.asarray(kernel)
    cdt = np.float32 if dtype == np.float32 else np.float64
    nd = img.ndim
    ks = kernel.shape
    # image must be known on [ -(k-1-c), n-1+c ]  with c = k//2  (taps t in [-c, k-1-c])
    lo = [k - 1 - (k // 2) for k in ks]   # needed before index 0: x - t with t up to k-1-c
    hi = [k // 2 for k in ks]             # needed after n-1:      x - t with t down to -c
    padded = _pad_volume(img.astype(cdt), list(zip(lo, hi)), ext, const)
    full = [padded.shape[d] for d in range(nd)]                      # n + k - 1
    fshape = [sfft.next_fast_len(s, real=(d == nd - 1)) for d, s in enumerate(full)]
    F = sfft.rfftn(padded, s=fshape, workers=workers)
    K = sfft.rfftn(kernel.astype(cdt), s=fshape, workers=workers)
    F *= K
    res = sfft.irfftn(F, s=fshape, workers=workers)
    # linear conv of padded (origin shifted by lo) with kernel (origin at index 0 == tap -c):
    # full[i] = sum_j padded[i - j] K[j];  y(x) = sum_j img_ext(x + c - j) K[j] = full[x + c + lo]
    sl = tuple(slice(lo[d] + ks[d] // 2, lo[d] + ks[d] // 2 + img.shape[d]) for d in range(nd))
    return np.ascontiguousarray(res[sl]).astype(dtype)


def circular_convolve(img: np.ndarray, kernel: np.ndarray, dtype=np.float32, workers: int = _WORKERS) -> np.ndarray:
    """Legacy GPU semantics (convolution3DfftCUDAInPlace, M/process/cuda/CUDAFourierConvolution.java:28-32
    as called from ComputeBlockSeqThreadCUDA.java:171-208): circular convolution at the image size,
    kernel zero-padded with its centre floor(k/2) moved to the origin."""
    img = np.asarray(img)
    cdt = np.float32 if dtype == np.float32 else np.float64
    kpad = np.zeros(img.shape, dtype=cdt)
    kpad[tuple(slice(0, k) for k in kernel.shape)] = kernel
    kpad = np.roll(kpad, [-(k // 2) for k in kernel.shape], axis=tuple(range(img.ndim)))
    F = sfft.rfftn(img.astype(cdt), workers=workers)
    F *= sfft.rfftn(kpad, workers=workers)
    return sfft.irfftn(F, s=img.shape, workers=workers).astype(dtype)


def direct_convolve(img: np.ndarray, kernel: np.ndarray, ext: str = "mirror", const: float = 1.0) -> np.ndarray:
    """O(N k^3) float64 direct evaluation of the same formula -- for tiny known-answer tests only."""
    img = np.asarray(img, dtype=np.float64)
    kernel = np.asarray(kernel, dtype=np.float64)
    ks = kernel.shape
    lo = [k - 1 - (k // 2) for k in ks]
    hi = [k // 2 for k in ks]
    p = _pad_volume(img, list(zip(lo, hi)), ext, const)
    out = np.zeros_like(img)
    for jz in range(ks[0]):
        for jy in range(ks[1]):
            for jx in range(ks[2]):
                # y(x) = sum_j img_ext(x + c - j) K[j];  padded index = x + c - j + lo
                oz = lo[0] + ks[0] // 2 - jz
                oy = lo[1] + ks[1] // 2 - jy
                ox = lo[2] + ks[2] // 2 - jx
                out += kernel[jz, jy, jx] * p[oz:oz + img.shape[0], oy:oy + img.shape[1], ox:ox + img.shape[2]]
    return out


# --------------------------------------------------------------------------------------
# DeconViewPSF.init   (M/process/deconvolution/DeconViewPSF.java:119-254)
# --------------------------------------------------------------------------------------
def derive_kernels(psfs: Sequence[np.ndarray], psf_type: int, quirk_threads=REFERENCE_THREADS,
                   dtype=np.float32) -> Tuple[List[np.ndarray], List[np.ndarray]]:
    """Returns (kernel1[], kernel2[]) exactly in the order DeconViews calls psf.init
    (M/process/deconvolution/DeconViews.java:69-70): view v's kernel1 is normalised at the
    start of its own init, so while building view v's compound kernel the kernel1 of views
    w < v are already normalised and those of w > v are not yet (Quirk B; harmless because of
    the final normalisation, but restated literally).

    PSF-derivation convolutions are zero-extended, output = size of the first operand
    (DeconViewPSF.java:152-178,215-225)."""
    k1 = [np.array(p, dtype=np.float32, copy=True) for p in psfs]
    V = len(k1)
    k2: List[Optional[np.ndarray]] = [None] * V
    for v in range(V):
        k1[v] = norm_to_sum1(k1[v], quirk_threads)                                    # :125
        if V == 1 or psf_type == INDEPENDENT:                                          # :127-131
            k2[v] = compute_inverted_kernel(k1[v])
        elif psf_type == EFFICIENT_BAYESIAN:                                           # :132-195
            tmp = compute_inverted_kernel(k1[v].copy())
            for w in range(V):
                if w == v:
                    continue
                inp = compute_inverted_kernel(k1[v])
                out = fft_convolve(inp, k1[w], ext="zero", dtype=dtype)
                out = fft_convolve(out, compute_inverted_kernel(k1[w]), ext="zero", dtype=dtype)
                tmp = (out.astype(np.float32) * tmp).astype(np.float32)
            k2[v] = norm_to_sum1(tmp, quirk_threads)
        elif psf_type == OPTIMIZATION_I:                                               # :196-242
            tmp = k1[v].copy()
            for w in range(V):
                if w == v:
                    continue
                out = fft_convolve(k1[v], compute_inverted_kernel(k1[w]), ext="zero", dtype=dtype)
                tmp = (out.astype(np.float32) * tmp).astype(np.float32)
            tmp = norm_to_sum1(tmp, quirk_threads)
            k2[v] = compute_inverted_kernel(tmp)
        else:                                                                          # OPTIMIZATION_II :243-253
            e = compute_exponential_kernel(k1[v], V)
            e = norm_to_sum1(e, quirk_threads)
            k2[v] = compute_inverted_kernel(e)
    return k1, [np.ascontiguousarray(k) for k in k2]


# --------------------------------------------------------------------------------------
# DeconvolutionMethods   (M/process/deconvolution/iteration/sequential/DeconvolutionMethods.java)
# --------------------------------------------------------------------------------------
def compute_quotient(psi_blurred: np.ndarray, observed: np.ndarray) -> np.ndarray:
    """q = img > 0 ? img / blurred : 1  in float32 (:46-98). No guard on blurred == 0."""
    b = np.asarray(psi_blurred, dtype=np.float32)
    o = np.asarray(observed, dtype=np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        q = (o / b).astype(np.float32)
    return np.where(o > 0, q, np.float32(1.0)).astype(np.float32)


def _tikhonov(value: np.ndarray, lam: float) -> np.ndarray:
    """(sqrt(1 + 2*lambda*value) - 1) / lambda  in float64 (:421)."""
    return (np.sqrt(1.0 + 2.0 * lam * value) - 1.0) / lam


def compute_next_value(last_psi, integral, weight, lam: float, min_intensity, max_intensity) -> np.ndarray:
    """computeNextValue (:320-358): float32 except the Tikhonov term."""
    f32 = np.float32
    last = np.asarray(last_psi, dtype=f32)
    with np.errstate(all="ignore"):
        value = (last * np.asarray(integral, dtype=f32)).astype(f32)
        if lam > 0:
            # (float)tikhonov( value / maxIntensity, lambda ) * maxIntensity : the division is f32,
            # promoted to double for tikhonov(), cast back to f32, multiplied in f32  (:337)
            ratio = (value / f32(max_intensity)).astype(f32).astype(np.float64)
            adjusted_pos = (_tikhonov(ratio, float(lam)).astype(f32) * f32(max_intensity)).astype(f32)
        else:
            adjusted_pos = value
        adjusted = np.where(value > 0, adjusted_pos, f32(min_intensity)).astype(f32)
        nxt = np.where(np.isnan(adjusted), f32(min_intensity), np.maximum(f32(min_intensity), adjusted)).astype(f32)
        # return lastPsiValue + ( ( nextPsiValue - lastPsiValue ) * weight )   (:357)
        return (last + ((nxt - last).astype(f32) * np.asarray(weight, dtype=f32)).astype(f32)).astype(f32)


def compute_next_value_mul(last_psi, integrals: Sequence[np.ndarray], weights: Sequence[np.ndarray], lam: float,
                           min_intensity, max_intensity) -> np.ndarray:
    """computeNextValueMul (mul/DeconvolutionMethods.java:370-419): geometric mean in float64."""
    f32 = np.float32
    last = np.asarray(last_psi, dtype=f32)
    V = len(weights)
    prod = np.ones(last.shape, dtype=np.float64)
    sumw = np.zeros(last.shape, dtype=np.float64)
    for i in range(V):
        prod = prod * np.asarray(integrals[i], dtype=f32).astype(np.float64)
        sumw = sumw + np.asarray(weights[i], dtype=f32).astype(np.float64)
    with np.errstate(all="ignore"):
        prod = np.power(prod, 1.0 / V)
        sumw = np.minimum(1.0, sumw)
        value = (last * prod.astype(f32)).astype(f32)
        if lam > 0:
            ratio = (value / f32(max_intensity)).astype(f32).astype(np.float64)
            adjusted_pos = (_tikhonov(ratio, float(lam)).astype(f32) * f32(max_intensity)).astype(f32)
        else:
            adjusted_pos = value
        adjusted = np.where(value > 0, adjusted_pos, f32(min_intensity)).astype(f32)
        nxt = np.where(np.isnan(adjusted), f32(min_intensity), np.maximum(f32(min_intensity), adjusted)).astype(f32)
        return (last + ((nxt - last).astype(f32) * sumw.astype(f32)).astype(f32)).astype(f32)


def iteration_statistics(last_psi: np.ndarray, next_psi: np.ndarray) -> Tuple[float, float]:
    """sumChange (f64 sum of signed f32 change) and maxChange = max(-1, max change)  (:114-115,147-149,308)."""
    change = (np.asarray(next_psi, dtype=np.float32) - np.asarray(last_psi, dtype=np.float32)).astype(np.float32)
    if change.size == 0:
        return 0.0, -1.0
    with np.errstate(all="ignore"):
        mx = float(np.fmax.reduce(change.ravel().astype(np.float64), initial=-1.0))  # Math.max ignores nothing, NaN-propagation differs; fmax is the safe variant
    return float(change.sum(dtype=np.float64)), max(-1.0, mx)


# --------------------------------------------------------------------------------------
# Block / BlockGeneratorFixedSizePrecise / BlockSorter   (M/process/cuda/*.java)
# --------------------------------------------------------------------------------------
@dataclass
class Block:
    """Dimension order (x, y, z) as in the reference (d = 0 is x)."""
    block_size: Tuple[int, ...]
    offset: Tuple[int, ...]
    effective_size: Tuple[int, ...]
    effective_offset: Tuple[int, ...]
    effective_local_offset: Tuple[int, ...]

    def min(self, d: int) -> int:
        return self.offset[d]

    def copy_block(self, source: np.ndarray, ext: str = "mirror", const: float = 0.0) -> np.ndarray:
        """Block.copyBlock (Block.java:158-197,277-315): cut [offset, offset+blockSize) from the
        out-of-bounds-extended source. source is [z,y,x]; returns [z,y,x] of block_size."""
        bs = self.block_size[::-1]
        off = self.offset[::-1]
        n = source.shape
        lo = [max(0, -off[d]) for d in range(3)]
        hi = [max(0, off[d] + bs[d] - n[d]) for d in range(3)]
        p = _pad_volume(np.asarray(source), list(zip(lo, hi)), ext, const)
        sl = tuple(slice(off[d] + lo[d], off[d] + lo[d] + bs[d]) for d in range(3))
        return np.ascontiguousarray(p[sl])

    def paste_block(self, target: np.ndarray, block: np.ndarray) -> None:
        """Block.pasteBlock (Block.java:199-239,357-403): effective region only."""
        es = self.effective_size[::-1]
        eo = self.effective_offset[::-1]
        el = self.effective_local_offset[::-1]
        src = block[tuple(slice(el[d], el[d] + es[d]) for d in range(3))]
        target[tuple(slice(eo[d], eo[d] + es[d]) for d in range(3))] = src


def divide_into_blocks(img_size: Sequence[int], block_size: Sequence[int], kernel_size: Sequence[int]) -> Optional[List[Block]]:
    """BlockGeneratorFixedSizePrecise.divideIntoBlocks (BlockGeneratorFixedSizePrecise.java:59-131).
    All arguments in (x, y, z) order; kernel_size is the *total* kernel (DeconView passes 2k-1).
    Block iteration order = LocalizingZeroMinIntervalIterator: x fastest."""
    nd = len(img_size)
    eff_general = [block_size[d] - kernel_size[d] + 1 for d in range(nd)]
    if any(e <= 0 for e in eff_general):
        return None
    eff_local = [kernel_size[d] // 2 for d in range(nd)]
    num_blocks = [img_size[d] // eff_general[d] + (1 if img_size[d] % eff_general[d] != 0 else 0) for d in range(nd)]
    blocks = []
    total = int(np.prod(num_blocks))
    for idx in range(total):
        cur = []
        r = idx
        for d in range(nd):
            cur.append(r % num_blocks[d])
            r //= num_blocks[d]
        eff_off = [cur[d] * eff_general[d] for d in range(nd)]
        off = [eff_off[d] - kernel_size[d] // 2 for d in range(nd)]
        eff_size = list(eff_general)
        for d in range(nd):
            if eff_off[d] + eff_size[d] > img_size[d]:
                eff_size[d] = img_size[d] - eff_off[d]
        blocks.append(Block(tuple(block_size), tuple(off), tuple(eff_size), tuple(eff_off), tuple(eff_local)))
    return blocks


def sort_blocks_by_smallest_footprint(blocks: List[Block], psi_dims: Sequence[int], min_required_blocks: int = 1) -> List[List[Block]]:
    """BlockSorter.sortBlocksBySmallestFootprint (BlockSorter.java:55-143), psi_dims in (x,y,z).
    Note the HashMap<size, dim> at :76-88: when two dimensions have the same orthogonal block count
    the *later* dimension overwrites the earlier one."""
    n = len(psi_dims)
    eff = blocks[0].effective_size
    num_blocks = []
    for d in range(n):
        nb = psi_dims[d] // eff[d]
        if psi_dims[d] % eff[d] != 0:
            nb += 1
        num_blocks.append(int(nb))
    size_to_dim = {}
    sizes = []
    for d in range(n):
        size = 1
        for e in range(n):
            if e != d:
                size *= num_blocks[e]
        sizes.append(size)
        size_to_dim[size] = d
    sizes.sort()
    min_dim = -1
    for i in range(n):
        if min_dim != -1:
            break
        if sizes[i] >= min_required_blocks or (i == n - 1 and min_dim == -1):
            min_dim = size_to_dim[sizes[i]]
    out: List[List[Block]] = []
    min_offset = blocks[0].offset
    total = 0
    for i in range(num_blocks[min_dim]):
        offset = min_offset[min_dim] + i * eff[min_dim]
        layer = [b for b in blocks if b.min(min_dim) == offset]
        total += len(layer)
        out.append(layer)
    if total != len(blocks):
        return [list(blocks)]
    return out


def block_contains_content(block: Block, weight: np.ndarray) -> bool:
    """DeconView.blockContainsContent (DeconView.java:236-274): any non-zero weight in the whole
    block (halo included), weight zero-extended."""
    return bool(np.any(block.copy_block(weight, ext="zero") != 0.0))


# --------------------------------------------------------------------------------------
# views / driver   (DeconView.java:118-184, MultiViewDeconvolutionSeq.java:58-180)
# --------------------------------------------------------------------------------------
@dataclass
class OracleView:
    image: np.ndarray      # [z,y,x] float32, 0 outside coverage
    weight: np.ndarray     # [z,y,x] float32
    kernel1: np.ndarray    # [z,y,x] float32 (normalised)
    kernel2: np.ndarray
    max_intensity: float = 1.0


def _conv1(psi, k1, dtype):   # ComputeBlockSeqThreadCPU.convolve1 (:171-189): mirror-single
    return fft_convolve(psi, k1, ext="mirror", dtype=dtype)


def _conv2(ratio, k2, dtype):  # ComputeBlockSeqThreadCPU.convolve2 (:191-209): constant 1
    return fft_convolve(ratio, k2, ext="const", const=1.0, dtype=dtype)


def view_update_whole(psi: np.ndarray, view: OracleView, lam: float, min_value=MIN_VALUE, dtype=np.float32):
    """One view update on the whole volume == ComputeBlockSeqThreadCPU.runIteration (:79-169) with a
    single block covering everything. Returns (psi_next, sumChange, maxChange).
    With dtype=float64 only the two FFT convolutions run in double; the pointwise stages keep the
    reference's float32 arithmetic (they are part of the specification, not of FFT rounding)."""
    blur = _conv1(psi, view.kernel1, dtype)
    ratio = compute_quotient(blur.astype(np.float32), view.image)
    integ = _conv2(ratio, view.kernel2, dtype)
    nxt = compute_next_value(psi, integ.astype(np.float32), view.weight, lam, min_value, view.max_intensity)
    s, m = iteration_statistics(psi, nxt)
    return nxt, s, m


def view_update_blocked(psi: np.ndarray, view: OracleView, block_size_xyz: Sequence[int], lam: float,
                        min_value=MIN_VALUE, dtype=np.float32, gpu_style: bool = False,
                        filter_blocks: bool = True, min_required_blocks: int = 1):
    """MultiViewDeconvolutionSeq.runNextIteration for ONE view, literally: blocks with halo 2k-1,
    copy-in with mirror OOB, img/weight zero OOB, delayed write-back by batch
    (MultiViewDeconvolutionSeq.java:69-176).  gpu_style=True uses the legacy CUDA semantics for the two
    convolutions (circular at block size, ComputeBlockSeqThreadCUDA.java:171-208)."""
    kz, ky, kx = view.kernel1.shape
    img_size = psi.shape[::-1]
    ksz = (2 * kx - 1, 2 * ky - 1, 2 * kz - 1)                      # DeconView.java:155-157
    blocks = divide_into_blocks(img_size, block_size_xyz, ksz)
    if blocks is None:
        raise ValueError("block smaller than kernel")
    batches = sort_blocks_by_smallest_footprint(blocks, img_size, min_required_blocks)
    if filter_blocks:                                               # DeconView.java:176-182,204-234
        batches = [[b for b in batch if block_contains_content(b, view.weight)] for batch in batches]
        batches = [b for b in batches if len(b) > 0]
    total_blocks = sum(len(b) for b in batches)
    psi = psi.copy()
    sum_change, max_change = 0.0, -1.0
    prev_q: List[Tuple[Block, np.ndarray]] = []
    for batch in batches:
        cur_q: List[Tuple[Block, np.ndarray]] = []
        for blk in batch:
            pb = blk.copy_block(psi, ext="mirror")
            ib = blk.copy_block(view.image, ext="zero")
            wb = blk.copy_block(view.weight, ext="zero")
            if gpu_style:
                blur = circular_convolve(pb, view.kernel1, dtype)
                ratio = compute_quotient(blur.astype(np.float32), ib)
                integ = circular_convolve(ratio, view.kernel2, dtype)
            else:
                blur = _conv1(pb, view.kernel1, dtype)
                ratio = compute_quotient(blur.astype(np.float32), ib)
                integ = _conv2(ratio, view.kernel2, dtype)
            nb = compute_next_value(pb, integ.astype(np.float32), wb, lam, min_value, view.max_intensity)
            s, m = iteration_statistics(pb, nb)     # stats are over the WHOLE block incl. halo (:129-157)
            sum_change += s
            max_change = max(max_change, m)
            if total_blocks == 1:
                blk.paste_block(psi, nb)
            else:
                cur_q.append((blk, nb))
        for blk, nb in prev_q:                       # write back the previous batch (:153-157)
            blk.paste_block(psi, nb)
        prev_q = cur_q
    for blk, nb in prev_q:
        blk.paste_block(psi, nb)
    return psi, sum_change, max_change


def run_iterations_seq(psi0: np.ndarray, views: Sequence[OracleView], num_iterations: int, lam: float,
                       min_value=MIN_VALUE, dtype=np.float32, block_size_xyz: Optional[Sequence[int]] = None,
                       gpu_style: bool = False, callback=None):
    """MultiViewDeconvolution.runIterations + MultiViewDeconvolutionSeq.runNextIteration
    (OSEM: psi is updated after every view)."""
    psi = np.asarray(psi0, dtype=np.float32).copy()
    stats = []
    for it in range(num_iterations):
        for v, view in enumerate(views):
            if block_size_xyz is None:
                psi, s, m = view_update_whole(psi, view, lam, min_value, dtype)
            else:
                psi, s, m = view_update_blocked(psi, view, block_size_xyz, lam, min_value, dtype, gpu_style)
            stats.append((it, v, s, m))
            if callback is not None:
                callback(it, v, psi, s, m)
    return psi, stats


def iteration_mul_whole(psi: np.ndarray, views: Sequence[OracleView], lam: float, min_value=MIN_VALUE, dtype=np.float32):
    """ComputeBlockMulThreadCPU.runIteration (mul/ComputeBlockMulThreadCPU.java:87-188) on one block
    covering the whole volume: all views from the same psi, geometric mean, max := mean of maxima."""
    integ, weights = [], []
    for view in views:
        blur = _conv1(psi, view.kernel1, dtype)
        ratio = compute_quotient(blur.astype(np.float32), view.image)
        integ.append(_conv2(ratio, view.kernel2, dtype).astype(np.float32))
        weights.append(view.weight)
    miv = 0.0
    for view in views:                      # double accumulation of Float values (:144-149)
        miv += float(np.float32(view.max_intensity))
    miv = np.float32(miv / float(len(views)))
    nxt = compute_next_value_mul(psi, integ, weights, lam, min_value, miv)
    s, m = iteration_statistics(psi, nxt)
    return nxt, s, m


# --------------------------------------------------------------------------------------
# weights: blending + normalisation
# --------------------------------------------------------------------------------------
def _blend_lut() -> np.ndarray:
    """BlendingRealRandomAccess static LUT (M/process/fusion/transformed/weights/BlendingRealRandomAccess.java:47-56),
    including the accumulating ``d = d + 0.001`` loop variable."""
    lut = np.zeros(1001, dtype=np.float64)
    d = 0.0
    while d <= 1.0001:
        idx = int(d * 1000.0 + 0.5)
        if idx <= 1000:
            lut[idx] = (math.cos((1 - d) * math.pi) + 1) / 2
        d = d + 0.001
    return lut


_BLEND_LUT = _blend_lut()


def blending_weight(dims_zyx: Sequence[int], box_min_xyz: Sequence[int], box_max_xyz: Sequence[int],
                    border: Sequence[float], blending: Sequence[float]) -> np.ndarray:
    """BlendingRealRandomAccess.computeWeight (:95-130) for an axis-aligned box [min, max] (inclusive,
    (x,y,z) order) evaluated on the integer grid of a [z,y,x] volume. float32 arithmetic as in the reference;
    the LUT multiply is float*double -> float per compound assignment."""
    f32 = np.float32
    nz, ny, nx = dims_zyx
    axes = [np.arange(nx, dtype=f32), np.arange(ny, dtype=f32), np.arange(nz, dtype=f32)]
    zero = [None, None, None]
    fac = [None, None, None]
    for d in range(3):
        mn = int(box_min_xyz[d])
        dim_minus1 = int(box_max_xyz[d]) - mn
        l = (axes[d] - f32(mn)).astype(f32)
        dist = np.minimum((l - f32(border[d])).astype(f32), (f32(dim_minus1) - l - f32(border[d])).astype(f32))
        zero[d] = dist <= 0
        rel = (dist / f32(blending[d])).astype(f32)
        with np.errstate(invalid="ignore"):
            idx = np.clip((rel.astype(np.float64) * 1000.0 + 0.5).astype(np.int64), 0, 1000)
        fac[d] = np.where(rel < 1, _BLEND_LUT[idx], 1.0)          # float64 factors
    wx, wy, wz = fac
    # minDistance (float) *= lookUp (double), in dimension order x, y, z
    w = np.ones((nz, ny, nx), dtype=f32)
    w = (w.astype(np.float64) * wx[None, None, :]).astype(f32)
    w = (w.astype(np.float64) * wy[None, :, None]).astype(f32)
    w = (w.astype(np.float64) * wz[:, None, None]).astype(f32)
    mask = zero[2][:, None, None] | zero[1][None, :, None] | zero[0][None, None, :]
    w[mask] = 0
    return w


def blending_weight_affine(dims_zyx: Sequence[int], bbox_offset_xyz: Sequence[int], inv_affine_row_packed: Sequence[float],
                           img_min_xyz: Sequence[int], img_max_xyz: Sequence[int], border: Sequence[float], blending: Sequence[float]) -> np.ndarray:
    """TransformWeight.transformBlending (M/process/fusion/transformed/TransformWeight.java:82-139): for every voxel of the zero-min fused
    volume, (voxel + offset) goes through the inverse affine in double, evaluated left to right (TransformedRasteredRandomAccess.applyInverse,
    M/process/fusion/transformed/weights/TransformedRasteredRandomAccess.java:96-114), is cast to float and weighted by
    BlendingRealRandomAccess.computeWeight (:95-130) of the view's image interval."""
    f32 = np.float32
    nz, ny, nx = dims_zyx
    im = np.asarray(inv_affine_row_packed, dtype=np.float64).ravel()
    t2, t1, t0 = np.meshgrid(np.arange(nz, dtype=np.float64) + bbox_offset_xyz[2], np.arange(ny, dtype=np.float64) + bbox_offset_xyz[1],
                             np.arange(nx, dtype=np.float64) + bbox_offset_xyz[0], indexing="ij")
    loc = [(((t0 * im[4 * r] + t1 * im[4 * r + 1]) + t2 * im[4 * r + 2]) + im[4 * r + 3]).astype(f32) for r in range(3)]
    w = np.ones((nz, ny, nx), dtype=f32)
    zero = np.zeros((nz, ny, nx), dtype=bool)
    facs = []
    for d in range(3):
        mn = int(img_min_xyz[d])
        dim_minus1 = int(img_max_xyz[d]) - mn
        l = (loc[d] - f32(mn)).astype(f32)
        dist = np.minimum((l - f32(border[d])).astype(f32), ((f32(dim_minus1) - l).astype(f32) - f32(border[d])).astype(f32))
        zero |= dist <= 0
        rel = (dist / f32(blending[d])).astype(f32)
        with np.errstate(invalid="ignore"):
            idx = np.clip((rel.astype(np.float64) * 1000.0 + 0.5).astype(np.int64), 0, 1000)
        facs.append(np.where(rel < 1, _BLEND_LUT[idx], 1.0))
    for fac in facs:                       # minDistance (float) *= lookUp (double), dimension order x, y, z
        w = (w.astype(np.float64) * fac).astype(f32)
    w[zero] = 0
    return w


# -----------------------------------------------------------------------------------------------------------------------
# Input view materialisation: TransformView + FusedRandomAccess(AVG) + CombineWeights(SUM) of one group,
# ProcessInputImages.fuseGroups (M/process/deconvolution/util/ProcessInputImages.java:279-399).
# The n-linear / nearest-neighbour samplers live in imglib2 8.0.0 (pom.xml:110), which is not in /root/reference: restated from
# its published NLinearInterpolator3D (Gray-code corner order 000,100,110,010,011,111,101,001; every corner value is
# multiplied by its double weight product and rounded to float -- FloatType.mul(double) -- then added in float) and
# NearestNeighborInterpolator (Round: half away from zero; positions are positive here).  PARITY UNPINNED like the rest.
# -----------------------------------------------------------------------------------------------------------------------
def _world_to_raw(dims_zyx, bbox_min_xyz, inv_affine_row_packed):
    """s = position + offset (TransformedInputRandomAccess.java:63-66), t = applyInverse(s): double, left to right (:69)."""
    nz, ny, nx = dims_zyx
    im = np.asarray(inv_affine_row_packed, dtype=np.float64).ravel()
    s2, s1, s0 = np.meshgrid(np.arange(nz, dtype=np.float64) + bbox_min_xyz[2], np.arange(ny, dtype=np.float64) + bbox_min_xyz[1],
                             np.arange(nx, dtype=np.float64) + bbox_min_xyz[0], indexing="ij")
    return [((s0 * im[4 * r] + s1 * im[4 * r + 1]) + s2 * im[4 * r + 2]) + im[4 * r + 3] for r in range(3)]


def nlinear3(raw: np.ndarray, t: Sequence[np.ndarray], extend_zero: bool = False) -> np.ndarray:
    """imglib2 NLinearInterpolator3D.get() for FloatType at real positions t = [tx, ty, tz] (see the section header).
    extend_zero: Views.extendZero (PSFExtraction.transform) instead of requiring the 8 corners inside."""
    f32, f64 = np.float32, np.float64
    rz, ry, rx = raw.shape
    fl = [np.floor(c) for c in t]
    w = [c - f for c, f in zip(t, fl)]
    wi = [1.0 - a for a in w]
    base = [f.astype(np.int64) for f in fl]

    def corner(dx, dy, dz):
        x, y, z = base[0] + dx, base[1] + dy, base[2] + dz
        ok = (x >= 0) & (x < rx) & (y >= 0) & (y < ry) & (z >= 0) & (z < rz)
        v = raw[np.clip(z, 0, rz - 1), np.clip(y, 0, ry - 1), np.clip(x, 0, rx - 1)].astype(f32)
        return np.where(ok, v, f32(0)) if extend_zero else v

    def term(dx, dy, dz):
        wx = w[0] if dx else wi[0]
        wy = w[1] if dy else wi[1]
        wz = w[2] if dz else wi[2]
        return (corner(dx, dy, dz).astype(f64) * ((wx * wy) * wz)).astype(f32)

    acc = term(0, 0, 0)
    for c in ((1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 1, 1), (1, 1, 1), (1, 0, 1), (0, 0, 1)):
        acc = (acc + term(*c)).astype(f32)
    return acc


def transform_view(raw: np.ndarray, inv_affine_row_packed, bbox_min_xyz, dims_zyx, interpolation: int = 1, has_min_value: bool = True,
                   min_value: float = MIN_VALUE_IMG, outside_value: float = OUTSIDE_VALUE_IMG) -> np.ndarray:
    """TransformView.transformView (M/process/fusion/transformed/TransformView.java:59-75) evaluated on the whole bounding box:
    TransformedInputRandomAccess.get (:60-82) with the strict inside test of AbstractTransformedIntervalRandomAccess.java:73-84 and
    getInsideValue = max(minValue, sample) (AbstractTransformedImgRandomAccess.java:76-91)."""
    f32 = np.float32
    raw = np.asarray(raw, dtype=f32)
    rz, ry, rx = raw.shape
    t = _world_to_raw(dims_zyx, bbox_min_xyz, inv_affine_row_packed)
    inside = (t[0] > 0) & (t[1] > 0) & (t[2] > 0) & (t[0] < rx - 1) & (t[1] < ry - 1) & (t[2] < rz - 1)
    tc = [np.where(inside, c, 0.25) for c in t]                     # any valid position for the masked-out voxels
    if interpolation == 1:
        val = nlinear3(raw, tc)
    else:
        idx = [np.trunc(c + 0.5 * np.sign(c)).astype(np.int64) for c in tc]       # Util.roundToLong
        val = raw[np.clip(idx[2], 0, rz - 1), np.clip(idx[1], 0, ry - 1), np.clip(idx[0], 0, rx - 1)]
    if has_min_value:
        val = np.maximum(f32(min_value), val)
    return np.where(inside, val, f32(outside_value)).astype(f32)


def fuse_group(raws: Sequence[np.ndarray], inv_affines, bbox_min_xyz, dims_zyx, interpolation: int = 1,
               fusion_blending=None, decon_blending=None, min_value: float = MIN_VALUE_IMG, outside_value: float = OUTSIDE_VALUE_IMG):
    """ProcessInputImages.fuseGroups for ONE group (ProcessInputImages.java:307-393): image = FusedRandomAccess AVG
    (M/process/fusion/transformed/FusedRandomAccess.java:66-91: views with weight 0 are skipped, sums in double, 0 where nothing
    contributes) of the transformed views with the fusion blending weights (or 1), weight = CombineWeightsSumRandomAccess
    (weightcombination/CombineWeightsSumRandomAccess.java:40-51) of the deconvolution blending weights (or 1).
    *_blending: None or a list of (border_xyz, range_xyz) per view, already adjusted (FusionTools.adjustBlending)."""
    f32, f64 = np.float32, np.float64
    sum_i = np.zeros(dims_zyx, dtype=f64)
    sum_w = np.zeros(dims_zyx, dtype=f64)
    sum_d = np.zeros(dims_zyx, dtype=f64)
    for j, (raw, ia) in enumerate(zip(raws, inv_affines)):
        img = transform_view(raw, ia, bbox_min_xyz, dims_zyx, interpolation, True, min_value, outside_value)
        rz, ry, rx = np.asarray(raw).shape
        mx = (rx - 1, ry - 1, rz - 1)
        wf = np.ones(dims_zyx, dtype=f32) if fusion_blending is None else \
            blending_weight_affine(dims_zyx, bbox_min_xyz, ia, (0, 0, 0), mx, fusion_blending[j][0], fusion_blending[j][1])
        wd = np.ones(dims_zyx, dtype=f32) if decon_blending is None else \
            blending_weight_affine(dims_zyx, bbox_min_xyz, ia, (0, 0, 0), mx, decon_blending[j][0], decon_blending[j][1])
        nz = wf != 0
        sum_i = np.where(nz, sum_i + img.astype(f64) * wf.astype(f64), sum_i)
        sum_w = np.where(nz, sum_w + wf.astype(f64), sum_w)
        sum_d = sum_d + wd.astype(f64)
    with np.errstate(all="ignore"):
        fused = np.where(sum_w > 0, (sum_i / sum_w).astype(f32), f32(0)).astype(f32)
    return fused, sum_d.astype(f32)


# ---- PSF preparation: PSFPreparation.loadGroupTransformPSFs (M/process/deconvolution/util/PSFPreparation.java:41-89) ---------------
def psf_normalize_minmax(psf: np.ndarray) -> np.ndarray:
    """PSFExtraction.normalize (M/process/psf/PSFExtraction.java:453-473): (v - min) / (max - min) in double, stored as float."""
    p = np.asarray(psf, dtype=np.float32).astype(np.float64)
    return ((p - p.min()) / (p.max() - p.min())).astype(np.float32)


def _apply_affine(m, x, y, z):
    return [((x * m[4 * r] + y * m[4 * r + 1]) + z * m[4 * r + 2]) + m[4 * r + 3] for r in range(3)]


def psf_transformed_geometry(dims_xyz, affine_row_packed):
    """new size (odd per axis) and offset of PSFExtraction.transformPSF (:367-409): estimateBounds of the interval [0, dim-1]
    (imglib2-realtransform: min / max over the 8 transformed corners), newSize = (int)size + 1 made odd, offset = A(center) - newSize/2
    with center = dim / 2 (integer division)."""
    m = np.asarray(affine_row_packed, dtype=np.float64).ravel()
    pts = np.array([_apply_affine(m, float(cx), float(cy), float(cz))
                    for cz in (0, dims_xyz[2] - 1) for cy in (0, dims_xyz[1] - 1) for cx in (0, dims_xyz[0] - 1)])
    size = pts.max(axis=0) - pts.min(axis=0)
    new = [int(size[d]) + 1 for d in range(3)]
    new = [n + 1 if n % 2 == 0 else n for n in new]
    ctr = _apply_affine(m, float(dims_xyz[0] // 2), float(dims_xyz[1] // 2), float(dims_xyz[2] // 2))
    off = [ctr[d] - (new[d] // 2) for d in range(3)]
    return new, off


def psf_transform(psf: np.ndarray, affine_row_packed, inv_affine_row_packed) -> np.ndarray:
    """PSFExtraction.getTransformedNormalizedPSF (:182-193) = normalize + transformPSF -> transform (:411-451): n-linear sampling of
    the zero-extended PSF at inverse(model)(voxel + offset)."""
    p = psf_normalize_minmax(psf)
    dz, dy, dx = p.shape
    new, off = psf_transformed_geometry((dx, dy, dz), affine_row_packed)
    im = np.asarray(inv_affine_row_packed, dtype=np.float64).ravel()
    z, y, x = np.meshgrid(np.arange(new[2], dtype=np.float64) + off[2], np.arange(new[1], dtype=np.float64) + off[1],
                          np.arange(new[0], dtype=np.float64) + off[0], indexing="ij")
    t = _apply_affine(im, x, y, z)
    return nlinear3(p, t, extend_zero=True)


def psf_average(psfs: Sequence[np.ndarray], use_max: bool = False) -> np.ndarray:
    """PSFCombination.computeAverageImage (M/process/psf/PSFCombination.java:74-135): centre-aligned float accumulation into the
    min (or max) size of all inputs (zero-extended target: samples outside are dropped), divided by the count in double."""
    shapes = np.array([p.shape for p in psfs])
    size = shapes.max(axis=0) if use_max else shapes.min(axis=0)
    avg = np.zeros(tuple(size), dtype=np.float32)
    ac = [s // 2 for s in size]
    for p in psfs:
        p = np.asarray(p, dtype=np.float32)
        pc = [s // 2 for s in p.shape]
        sl_a, sl_p = [], []
        for d in range(3):
            lo = ac[d] - pc[d]                       # position of psf sample 0 in the average
            a0, a1 = max(0, lo), min(size[d], lo + p.shape[d])
            sl_a.append(slice(a0, a1)); sl_p.append(slice(a0 - lo, a1 - lo))
        avg[tuple(sl_a)] = (avg[tuple(sl_a)] + p[tuple(sl_p)]).astype(np.float32)
    return (avg.astype(np.float64) / float(len(psfs))).astype(np.float32)


def psf_make_same_size(psf: np.ndarray, size_zyx) -> np.ndarray:
    """PSFCombination.makeSameSize (:182-212): centred copy into the new size, padded with the minimum of the input."""
    p = np.asarray(psf, dtype=np.float32)
    out = np.full(tuple(size_zyx), p.min(), dtype=np.float32)
    idx = [np.arange(size_zyx[d]) - size_zyx[d] // 2 + p.shape[d] // 2 for d in range(3)]
    ok = [(i >= 0) & (i < p.shape[d]) for d, i in enumerate(idx)]
    zz, yy, xx = np.ix_(idx[0][ok[0]], idx[1][ok[1]], idx[2][ok[2]])
    oz, oy, ox = np.ix_(np.nonzero(ok[0])[0], np.nonzero(ok[1])[0], np.nonzero(ok[2])[0])
    out[oz, oy, ox] = p[zz, yy, xx]
    return out


def smooth_weights(w: np.ndarray, sumw: np.ndarray, max_diff_range=MAX_DIFF_RANGE, scaling_range=SCALING_RANGE) -> np.ndarray:
    """NormalizingRandomAccess.smoothWeights (NormalizingRandomAccess.java:183-201)."""
    f32 = np.float32
    w = np.asarray(w, dtype=f32)
    with np.errstate(all="ignore"):
        ideal = (w.astype(np.float64) / sumw).astype(f32)
        diff = (w - ideal).astype(f32)
        y = np.maximum(f32(0), ((f32(max_diff_range) - np.abs(diff)).astype(f32) * (f32(1.0) / f32(max_diff_range))).astype(f32))
        scale = ((y * w).astype(f32) * f32(scaling_range)).astype(f32)
        res = (np.minimum(w, ideal) - scale).astype(f32)
    return np.where(sumw <= 0, f32(0), res).astype(f32)


def normalize_weights(raw: Sequence[np.ndarray], osem_speedup: float = 1.0, additional_smooth: bool = False) -> List[np.ndarray]:
    """NormalizingRandomAccess.get (NormalizingRandomAccess.java:75-109) for every view."""
    f32 = np.float32
    u = [np.minimum(1.0, np.asarray(r, dtype=f32).astype(np.float64)) for r in raw]
    sumw = np.zeros(u[0].shape, dtype=np.float64)
    for x in u:
        sumw = sumw + x
    out = []
    for x in u:
        my = x.astype(f32)
        if additional_smooth:
            v = smooth_weights(my, sumw).astype(np.float64)
        else:
            with np.errstate(all="ignore"):
                hard = (my.astype(np.float64) / sumw).astype(f32).astype(np.float64)   # hardWeights returns float
            v = np.where(sumw > 1, hard, my.astype(np.float64))
        out.append(np.minimum(1.0, v * float(osem_speedup)).astype(f32))
    return out


# --------------------------------------------------------------------------------------
# PsiInit
# --------------------------------------------------------------------------------------
def gauss3_halfkernel(sigma: float) -> np.ndarray:
    """ASSUMPTION (third-party, not in tree): net.imglib2.algorithm.gauss3.Gauss3.halfkernel with
    size = max(2, (int)(3*sigma + 0.5) + 1), ImageJ-style smoothEdge taper, normalised so the full
    symmetric kernel sums to 1."""
    size = max(2, int(3 * sigma + 0.5) + 1)
    k = np.zeros(size, dtype=np.float64)
    k[0] = 1.0
    for x in range(1, size):
        k[x] = math.exp(-(x * x) / (2 * sigma * sigma))
    if size > 3:
        sqrt_slope = float("inf")
        r = size
        while r > size // 2:
            r -= 1
            a = math.sqrt(k[r]) / (size - r)
            if a < sqrt_slope:
                sqrt_slope = a
            else:
                break
        for r1 in range(r + 2, size):
            k[r1] = (size - r1) * (size - r1) * sqrt_slope * sqrt_slope
    s = 2 * (0.5 * k[0] + k[1:].sum())
    return k / s


def gauss3_mirror(vol: np.ndarray, sigma: float) -> np.ndarray:
    """Separable Gauss (x, then y, then z) with mirror-single extension, float32 storage between axes,
    float64 accumulation (ASSUMPTION, see gauss3_halfkernel)."""
    half = gauss3_halfkernel(sigma)
    full = np.concatenate([half[:0:-1], half])
    r = len(half) - 1
    out = np.asarray(vol, dtype=np.float32)
    for axis in (2, 1, 0):
        n = out.shape[axis]
        pads = [(0, 0)] * 3
        pads[axis] = (r, r)
        p = np.pad(out.astype(np.float64), pads, mode="reflect") if n > 1 else np.pad(out.astype(np.float64), pads, mode="edge")
        acc = np.zeros(out.shape, dtype=np.float64)
        for j, c in enumerate(full):
            sl = [slice(None)] * 3
            sl[axis] = slice(j, j + n)
            acc += c * p[tuple(sl)]
        out = acc.astype(np.float32)
    return out


def psi_init_fused_stats(images: Sequence[np.ndarray], weights: Sequence[np.ndarray]):
    """FusedNonZeroRandomAccess.get (M/process/deconvolution/util/FusedNonZeroRandomAccess.java:57-96) over
    the whole volume + the aggregation of PsiInitBlurredFused.runInitialization (:76-101).
    Returns (fused float32 volume, max[] float32, avg float64 or None when no voxel is covered)."""
    shape = images[0].shape
    sum_i = np.zeros(shape, dtype=np.float64)
    sum_w = np.zeros(shape, dtype=np.float64)
    ssum = np.zeros(shape, dtype=np.float64)
    count = np.zeros(shape, dtype=np.int64)
    mx = np.zeros(len(images), dtype=np.float32)
    for j, (im, w) in enumerate(zip(images, weights)):
        inten = np.asarray(im, dtype=np.float32).astype(np.float64)
        pos = inten > 0
        wt = np.asarray(w, dtype=np.float32).astype(np.float64)
        sum_i += np.where(pos, inten * wt, 0.0)
        sum_w += np.where(pos, wt, 0.0)
        ssum += np.where(pos, inten, 0.0)
        count += pos
        if pos.any():
            mx[j] = max(mx[j], np.float32(inten[pos].max()))
    with np.errstate(all="ignore"):
        fused = np.where(sum_w > 0, (sum_i / sum_w), 0.0).astype(np.float32)
    covered = count > 0
    n_cov = int(covered.sum())
    if n_cov == 0:
        return fused, mx, None
    avg = float((ssum[covered] / count[covered]).sum(dtype=np.float64) / n_cov)
    if math.isnan(avg):
        avg = 1.0
    return fused, mx, avg


def psi_init_blurred_fused(images, weights, sigma: float = 5.0):
    """PsiInitBlurredFused.runInitialization (M/process/deconvolution/init/PsiInitBlurredFused.java:63-127)."""
    fused, mx, avg = psi_init_fused_stats(images, weights)
    if avg is None:
        return None, mx, None
    return gauss3_mirror(fused, sigma), mx, avg


def psi_init_avg_precise(images, set_img_to_avg: bool = True, psi: Optional[np.ndarray] = None):
    """PsiInitAvgPrecise (init/PsiInitAvgPrecise.java:52-112, PsiInitAvgPreciseThread.java:127-155)."""
    shape = images[0].shape
    ssum = np.zeros(shape, dtype=np.float64)
    count = np.zeros(shape, dtype=np.int64)
    mx = np.zeros(len(images), dtype=np.float32)
    for j, im in enumerate(images):
        i32 = np.asarray(im, dtype=np.float32)
        pos = i32 > 0
        ssum += np.where(pos, i32.astype(np.float64), 0.0)
        count += pos
        if pos.any():
            mx[j] = i32[pos].max()
    covered = count > 0
    n_cov = int(covered.sum())
    with np.errstate(all="ignore"):
        avg = float((ssum[covered] / count[covered]).sum(dtype=np.float64)) / n_cov if n_cov else float("nan")
    if math.isnan(avg):
        avg = 1.0
    out = np.full(shape, np.float32(avg), dtype=np.float32) if set_img_to_avg else psi
    return out, mx, avg


def psi_init_avg_approx(images, set_img_to_avg: bool = True, psi: Optional[np.ndarray] = None):
    """PsiInitAvgApprox (init/PsiInitAvgApprox.java:47-99; PsiInitAvgApproxThread.java:58-85): min/max/mean of
    the central x-hyperslice, visited img.numDimensions() times (loop variable d unused, :69-82).
    getAvg() returns -1 because the field is shadowed by a local (:40,57,80) -- reproduced."""
    mx = np.zeros(len(images), dtype=np.float32)
    avg = 0.0
    for j, im in enumerate(images):
        a = np.asarray(im, dtype=np.float32)
        sl = a[:, :, a.shape[2] // 2].astype(np.float64)
        nd = a.ndim
        total = 0.0
        for _ in range(nd):
            total += float(sl.sum(dtype=np.float64))
        mean = total / float(nd * sl.size)
        avg += mean
        mx[j] = np.float32(sl.max())
    avg /= float(len(images))
    if math.isnan(avg):
        avg = 1.0
    shape = images[0].shape
    out = np.full(shape, np.float32(avg), dtype=np.float32) if set_img_to_avg else psi
    return out, mx, -1.0


# --------------------------------------------------------------------------------------
# seeded synthetic inputs -- counter-based SplitMix64 so C++/CUDA agree
# --------------------------------------------------------------------------------------
_M64 = (1 << 64) - 1


def splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = (np.asarray(x, dtype=np.uint64) + np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def rng_uniform(seed: int, stream: int, index: np.ndarray) -> np.ndarray:
    """U[0,1) double from (seed, stream, index)."""
    with np.errstate(over="ignore"):
        key = splitmix64(np.uint64(seed & _M64) * np.uint64(0x632BE59BD9B4E019) + np.uint64(stream))
        z = splitmix64(key + np.asarray(index, dtype=np.uint64))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def synth_psf(view: int, num_views: int, size_xyz=(25, 19, 25), sigma_xyz=(1.5, 1.5, 4.0), tilt_deg: Optional[float] = None) -> np.ndarray:
    """Anisotropic Gaussian rotated about the y axis by theta_v = v*180/V (or tilt_deg), divided by its sum. [z,y,x] float32."""
    kx, ky, kz = size_xyz
    theta = math.radians(view * 180.0 / num_views if tilt_deg is None else tilt_deg)
    x = np.arange(kx, dtype=np.float64) - kx // 2
    y = np.arange(ky, dtype=np.float64) - ky // 2
    z = np.arange(kz, dtype=np.float64) - kz // 2
    Z, Y, X = np.meshgrid(z, y, x, indexing="ij")
    c, s = math.cos(theta), math.sin(theta)
    xr = c * X + s * Z
    zr = -s * X + c * Z
    g = np.exp(-0.5 * ((xr / sigma_xyz[0]) ** 2 + (Y / sigma_xyz[1]) ** 2 + (zr / sigma_xyz[2]) ** 2))
    return (g / g.sum()).astype(np.float32)


def synth_truth(dims_zyx, seed: int, bead_density: int = 8192) -> np.ndarray:
    """background 100 + point beads (count = N / bead_density, integer positions, amplitude U[500,4000])."""
    nz, ny, nx = dims_zyx
    n = nz * ny * nx
    nb = max(1, n // bead_density)
    idx = np.arange(nb, dtype=np.uint64)
    px = np.minimum((rng_uniform(seed, 1, idx) * nx).astype(np.int64), nx - 1)
    py = np.minimum((rng_uniform(seed, 2, idx) * ny).astype(np.int64), ny - 1)
    pz = np.minimum((rng_uniform(seed, 3, idx) * nz).astype(np.int64), nz - 1)
    amp = (500.0 + 3500.0 * rng_uniform(seed, 4, idx)).astype(np.float32)
    truth = np.full(dims_zyx, 100.0, dtype=np.float32)
    np.add.at(truth, (pz, py, px), amp)
    return truth


def synth_coverage_box(dims_zyx, view: int):
    """full volume minus a slab of 1/8 of the extent on side (view mod 6): 0:-x 1:+x 2:-y 3:+y 4:-z 5:+z.
    Returns (min_xyz, max_xyz) inclusive."""
    nz, ny, nx = dims_zyx
    mn = [0, 0, 0]
    mx = [nx - 1, ny - 1, nz - 1]
    side = view % 6
    d = side // 2
    ext = (nx, ny, nz)[d]
    cut = ext // 8
    if side % 2 == 0:
        mn[d] = cut
    else:
        mx[d] = ext - 1 - cut
    return tuple(mn), tuple(mx)


@dataclass
class SynthDataset:
    dims_zyx: Tuple[int, int, int]
    psfs: List[np.ndarray]
    images: List[np.ndarray]
    weights: List[np.ndarray]
    truth: np.ndarray
    boxes: List[Tuple[Tuple[int, ...], Tuple[int, ...]]] = field(default_factory=list)


def make_synthetic(dims_zyx, num_views: int, seed: int, psf_size_xyz=(25, 19, 25), psf_sigma_xyz=(1.5, 1.5, 4.0),
                   blend_range: float = 12.0, blend_border: float = 0.0, bead_density: int = 8192,
                   tilt_step_deg: Optional[float] = None) -> SynthDataset:
    """SURVEY 8d generator: img_v = max(1, truth (*) PSF_v) inside view v's box, 0 outside; w_v = cosine blending
    of the box -> hard normalisation (osem 1)."""
    truth = synth_truth(dims_zyx, seed, bead_density)
    psfs, images, raw, boxes = [], [], [], []
    for v in range(num_views):
        tilt = None if tilt_step_deg is None else (v - num_views // 2) * tilt_step_deg
        psf = synth_psf(v, num_views, psf_size_xyz, psf_sigma_xyz, tilt)
        psfs.append(psf)
        blurred = fft_convolve(truth, psf, ext="mirror", dtype=np.float32)
        mn, mx = synth_coverage_box(dims_zyx, v)
        boxes.append((mn, mx))
        img = np.zeros(dims_zyx, dtype=np.float32)
        sl = (slice(mn[2], mx[2] + 1), slice(mn[1], mx[1] + 1), slice(mn[0], mx[0] + 1))
        img[sl] = np.maximum(MIN_VALUE_IMG, blurred[sl])
        images.append(img)
        raw.append(blending_weight(dims_zyx, mn, mx, (blend_border,) * 3, (blend_range,) * 3))
    weights = normalize_weights(raw, 1.0, False)
    return SynthDataset(tuple(dims_zyx), psfs, images, weights, truth, boxes)


def make_oracle_views(ds: SynthDataset, psf_type: int, quirk_threads=REFERENCE_THREADS):
    """kernels by PSFTYPE, psi0 / max[] by FUSED_BLURRED.  Returns (views, psi0, avg)."""
    k1, k2 = derive_kernels(ds.psfs, psf_type, quirk_threads)
    psi0, mx, avg = psi_init_blurred_fused(ds.images, ds.weights)
    views = [OracleView(ds.images[v], ds.weights[v], k1[v], k2[v], float(mx[v])) for v in range(len(ds.psfs))]
    return views, psi0, avg


def rel_l2(a: np.ndarray, b: np.ndarray) -> float:
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
