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
