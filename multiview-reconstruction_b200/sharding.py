"""z-slab sharding helpers for the multi-GPU path (one process per GPU, torch.distributed for the plumbing).

The fused volume is cut into `world` slabs along z (the slowest axis, so a halo of p planes is one contiguous region).
Rank r owns planes [lo, hi) and keeps the extended slab [z0, z1) = [lo - H, hi + H) clipped to the volume, where
H = (k1z - 1)/2 + (k2z - 1)/2 + ... = kz - 1 is what the two chained convolutions of a view update read beyond the owned planes
(the same halo rule as the reference's blocks, DeconView.java:155-157).  After every view update each rank sends its first / last H
owned planes of the new psi to its lower / upper neighbour (MultiViewDeconvolutionSeq updates psi after every view, so the exchange
is per view update, not per iteration).
"""
from __future__ import annotations

from typing import Tuple


def slab_range(nz: int, world: int, rank: int) -> Tuple[int, int]:
    return rank * nz // world, (rank + 1) * nz // world


def extended_range(lo: int, hi: int, nz: int, halo: int) -> Tuple[int, int]:
    return max(0, lo - halo), min(nz, hi + halo)


def exchange_halos(buf, plane: int, lo: int, hi: int, z0: int, halo: int, rank: int, world: int, dist) -> None:
    """buf: flat tensor of the extended slab (planes [z0, ...), `plane` elements each) holding the freshly updated psi on the
    owned planes.  Sends owned boundary planes to the neighbours and receives their planes into the halo (batched isend/irecv;
    works with the nccl and the gloo backend)."""
    if world == 1:
        return
    ops = []
    if rank > 0:
        ops.append(dist.P2POp(dist.isend, buf[(lo - z0) * plane:(lo - z0 + halo) * plane], rank - 1))
        ops.append(dist.P2POp(dist.irecv, buf[(lo - z0 - halo) * plane:(lo - z0) * plane], rank - 1))
    if rank < world - 1:
        ops.append(dist.P2POp(dist.isend, buf[(hi - z0 - halo) * plane:(hi - z0) * plane], rank + 1))
        ops.append(dist.P2POp(dist.irecv, buf[(hi - z0) * plane:(hi - z0 + halo) * plane], rank + 1))
    for req in dist.batch_isend_irecv(ops):
        req.wait()


# ---------------------------------------------------------------------------------------------------------------------
# 2-d (y x z) process grid
# ---------------------------------------------------------------------------------------------------------------------
def axis_cost(n: int, lo: int, hi: int, reach: int, lengths, scheme: int = 0) -> int:
    """FFT-box extent (tiles * T) the engine's planner needs to cover [lo, hi) of an axis of size n when each of the two chained
    convolutions reaches `reach` samples (mirrors plan_axis in csrc/engine.cpp; halo per interior side = 2 * reach with exchange
    scheme 0, reach with scheme 1)."""
    best = None
    two = scheme == 1 and (lo != 0 or hi != n)
    for T in lengths:
        pos, o, tiles, ok = lo, (lo - reach if (lo == 0 or two) else lo - 2 * reach), 0, True
        while pos < hi:
            vend = o + T - 2 * reach
            if (hi == n or two) and o + T >= hi + reach:
                vend = hi
            nxt = min(vend, hi)
            if nxt <= pos:
                ok = False
                break
            tiles += 1
            pos, o = nxt, nxt - 2 * reach
        if ok and (best is None or tiles * T < best):
            best = tiles * T
    if best is None:
        raise ValueError("no FFT length fits")
    return best


def axis_plan(n: int, lo: int, hi: int, reach: int, lengths, scheme: int = 0) -> Tuple[int, int]:
    """(FFT-box extent, tile length) of the cheapest plan of axis_cost"""
    cost = axis_cost(n, lo, hi, reach, lengths, scheme)
    for T in lengths:
        try:
            if axis_cost(n, lo, hi, reach, [T], scheme) == cost:
                return cost, T
        except ValueError:
            continue
    return cost, cost


def grid_for(world: int, ny: int, nz: int, reach_y: int, reach_z: int, lengths, scheme: int = 0) -> Tuple[int, int]:
    """(py, pz) with py * pz == world minimising the FFT-box volume of the slowest rank, evaluated with the library's real tile
    lengths (Lib.supported_fft_lengths()).  Among grids within 4 % of the smallest volume the one with the shortest transforms
    wins (measured on 8 B200: 4 x 2 boxes of 288 x 288 beat 8 x 1 boxes of 150 x 540 by 7 % although they are 2 % larger); at equal
    cost the larger py wins: the single-GPU plan already splits y into FFT tiles, so the first y split is free."""
    cands = []
    for py in range(1, world + 1):
        if world % py:
            continue
        pz = world // py
        if ny // py <= 4 * reach_y or nz // pz <= 4 * reach_z:
            continue
        py_plans = [axis_plan(ny, *slab_range(ny, py, r), reach_y, lengths, scheme) for r in range(py)]
        pz_plans = [axis_plan(nz, *slab_range(nz, pz, r), reach_z, lengths, scheme) for r in range(pz)]
        cy, cz = max(c for c, _ in py_plans), max(c for c, _ in pz_plans)
        tmax = max(max(t for _, t in py_plans), max(t for _, t in pz_plans))
        cands.append((cy * cz, tmax, -py, (py, pz)))
    if not cands:
        raise ValueError("volume too small for this many ranks")
    vmin = min(c[0] for c in cands)
    near = [c for c in cands if c[0] <= 1.04 * vmin]
    return min(near, key=lambda c: (c[1], c[0], c[2]))[3]


def exchange_halos_2d(buf3, own_y, loc_y, own_z, loc_z, halo_y, halo_z, ry, rz, py, pz, rank_of, dist) -> None:
    """buf3: tensor [nz_loc, ny_loc, nx] of the extended local box holding the updated psi on the owned box.
    Phase 1 exchanges y rows over the owned z planes, phase 2 exchanges z planes INCLUDING the freshly received y halos, so the
    corner regions are correct.  own_* = (lo, hi) global, loc_* = (start, size) of the local array along that axis."""
    ylo, yhi = own_y
    y0 = loc_y[0]
    zlo, zhi = own_z
    z0 = loc_z[0]
    if py > 1:
        zs = slice(zlo - z0, zhi - z0)
        ops, recvs = [], []
        if ry > 0:
            peer = rank_of(ry - 1, rz)
            send = buf3[zs, ylo - y0:ylo - y0 + halo_y, :].contiguous()
            rb = buf3.new_empty(send.shape)
            ops += [dist.P2POp(dist.isend, send, peer), dist.P2POp(dist.irecv, rb, peer)]
            recvs.append((rb, slice(ylo - y0 - halo_y, ylo - y0)))
        if ry < py - 1:
            peer = rank_of(ry + 1, rz)
            send = buf3[zs, yhi - y0 - halo_y:yhi - y0, :].contiguous()
            rb = buf3.new_empty(send.shape)
            ops += [dist.P2POp(dist.isend, send, peer), dist.P2POp(dist.irecv, rb, peer)]
            recvs.append((rb, slice(yhi - y0, yhi - y0 + halo_y)))
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        for rb, ys in recvs:
            buf3[zs, ys, :] = rb
    if pz > 1:
        ops = []
        if rz > 0:
            peer = rank_of(ry, rz - 1)
            ops += [dist.P2POp(dist.isend, buf3[zlo - z0:zlo - z0 + halo_z], peer), dist.P2POp(dist.irecv, buf3[zlo - z0 - halo_z:zlo - z0], peer)]
        if rz < pz - 1:
            peer = rank_of(ry, rz + 1)
            ops += [dist.P2POp(dist.isend, buf3[zhi - z0 - halo_z:zhi - z0], peer), dist.P2POp(dist.irecv, buf3[zhi - z0:zhi - z0 + halo_z], peer)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()


def exchange_box(t3, own_y, own_z, halo_y, halo_z, ry, rz, py, pz, rank_of, dist) -> None:
    """Generic two-phase exchange on a tensor [planes, rows, row_floats] (the body of an mvd_exchange_fn callback): own_* = (lo, hi)
    array indices of the own region, halo_* = (below, above) widths.  The neighbour below gets this box's first `above` own rows /
    planes and sends its last `below` ones; the neighbour above mirrors that.  y rows travel over the own planes only, then whole z
    planes including the fresh y halos (corners)."""
    y0, y1 = own_y
    z0, z1 = own_z
    hl, hu = halo_y
    if py > 1 and (hl or hu):
        zs = slice(z0, z1)
        ops, recvs = [], []
        for present, peer_ry, send_rows, recv_rows in ((ry > 0, ry - 1, slice(y0, y0 + hu), slice(y0 - hl, y0)),
                                                       (ry < py - 1, ry + 1, slice(y1 - hl, y1), slice(y1, y1 + hu))):
            if not present:
                continue
            peer = rank_of(peer_ry, rz)
            if send_rows.stop > send_rows.start:
                ops.append(dist.P2POp(dist.isend, t3[zs, send_rows, :].contiguous(), peer))
            if recv_rows.stop > recv_rows.start:
                rb = t3.new_empty((z1 - z0, recv_rows.stop - recv_rows.start, t3.shape[2]))
                ops.append(dist.P2POp(dist.irecv, rb, peer))
                recvs.append((rb, recv_rows))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        for rb, rows in recvs:
            t3[zs, rows, :] = rb
    hl, hu = halo_z
    if pz > 1 and (hl or hu):
        ops = []
        for present, peer_rz, send_pl, recv_pl in ((rz > 0, rz - 1, slice(z0, z0 + hu), slice(z0 - hl, z0)),
                                                   (rz < pz - 1, rz + 1, slice(z1 - hl, z1), slice(z1, z1 + hu))):
            if not present:
                continue
            peer = rank_of(ry, peer_rz)
            if send_pl.stop > send_pl.start:
                ops.append(dist.P2POp(dist.isend, t3[send_pl], peer))
            if recv_pl.stop > recv_pl.start:
                ops.append(dist.P2POp(dist.irecv, t3[recv_pl], peer))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()


def host_exchange_callback(ry, rz, py, pz, rank_of, dist, device=None):
    """An exchange callback for DeconViews.set_exchange_callback built on torch.distributed: wraps the library's buffer (host memory
    for the CPU emulator, device memory otherwise) as a tensor and runs exchange_box on it."""
    import ctypes

    import numpy as np
    import torch

    def cb(which, box):
        n = box.nplanes * box.nrows * box.row_floats
        if device is None:
            arr = np.ctypeslib.as_array(box.base, shape=(n,))
            t3 = torch.from_numpy(arr).view(box.nplanes, box.nrows, box.row_floats)
        else:
            from . import RawDeviceBuffer
            ptr = ctypes.cast(box.base, ctypes.c_void_p).value
            t3 = torch.as_tensor(RawDeviceBuffer(ptr, (box.nplanes, box.nrows, box.row_floats)), device=device)
        exchange_box(t3, (box.y0, box.y1), (box.z0, box.z1), (box.hy_lo, box.hy_hi), (box.hz_lo, box.hz_hi), ry, rz, py, pz, rank_of, dist)
        if device is not None:
            torch.cuda.synchronize(device)
    return cb


def host_reduce_callback(dist):
    """A reduce callback for DeconViews.set_reduce_callback built on torch.distributed (CPU tensors: works with gloo, and with
    nccl groups that have a gloo side group)."""
    import torch

    def cb(values, op):
        t = torch.from_numpy(values)
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == 1 else dist.ReduceOp.SUM)
    return cb
