"""Sharding of the enhancement path across the GPUs of one box.

  * slice batches (BASELINE.json configs 2, 4): a contiguous split of the batch
    dimension — `shard_range` — and no communication at all;
  * volumes (config 3): z-slabs.  The 3x3x3 median needs ONE neighbouring plane per
    interior slab face; `exchange_z_halos` posts one send/recv pair per face with
    torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests), and
    `median3d_clahe_slab` overlaps that exchange with the median of the slab's
    interior planes.  CLAHE is per slice and needs nothing further.

One process per GPU (torchrun); every function here is a no-op wrapper when
torch.distributed is not initialised (world size 1).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

__all__ = ["shard_range", "exchange_z_halos", "start_z_halo_exchange", "median3d_clahe_slab", "world_info"]


def world_info(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_range(total: int, world_size: int, rank: int):
    """Contiguous, balanced split: the first `total % world_size` ranks own one extra item."""
    if world_size <= 0 or not 0 <= rank < world_size:
        raise ValueError(f"invalid rank {rank} for world size {world_size}")
    base, rem = divmod(int(total), world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def _wire(t: torch.Tensor) -> torch.Tensor:
    """16-bit integer planes have no NCCL process-group datatype: ship the same bytes as uint8."""
    return t.view(torch.uint8)


class _HaloExchange:
    """Handle of an in-flight halo exchange; .wait() returns (halo_lo, halo_hi)."""

    def __init__(self, reqs, lo, hi):
        self._reqs, self._lo, self._hi = reqs, lo, hi

    def wait(self):
        for r in self._reqs:
            r.wait()
        self._reqs = []
        return self._lo, self._hi


def start_z_halo_exchange(slab: torch.Tensor, group=None) -> _HaloExchange:
    """Post the z-halo exchange of a (D, H, W) slab: the first / last owned plane go to the lower /
    upper neighbour rank, theirs come back.  Ranks that own an empty slab take no part and must not
    exist between non-empty ones (use shard_range).  Returns immediately."""
    rank, world = world_info(group)
    if slab.dim() != 3:
        raise ValueError("expected a (D, H, W) slab")
    if world == 1 or slab.shape[0] == 0:
        return _HaloExchange([], None, None)
    d, h, w = slab.shape
    lo = torch.empty((h, w), dtype=slab.dtype, device=slab.device) if rank > 0 else None
    hi = torch.empty((h, w), dtype=slab.dtype, device=slab.device) if rank < world - 1 else None
    first, last = slab[0].contiguous(), slab[d - 1].contiguous()
    ops = []
    if rank > 0:
        ops.append(dist.P2POp(dist.isend, _wire(first), dist.get_global_rank(group, rank - 1) if group else rank - 1, group))
        ops.append(dist.P2POp(dist.irecv, _wire(lo), dist.get_global_rank(group, rank - 1) if group else rank - 1, group))
    if rank < world - 1:
        ops.append(dist.P2POp(dist.isend, _wire(last), dist.get_global_rank(group, rank + 1) if group else rank + 1, group))
        ops.append(dist.P2POp(dist.irecv, _wire(hi), dist.get_global_rank(group, rank + 1) if group else rank + 1, group))
    reqs = dist.batch_isend_irecv(ops) if ops else []
    ex = _HaloExchange(reqs, lo, hi)
    ex._keep = (first, last)  # keep the send buffers alive until wait()
    return ex


def exchange_z_halos(slab: torch.Tensor, group=None):
    """Blocking form of start_z_halo_exchange: returns (halo_lo, halo_hi), None at a volume face."""
    return start_z_halo_exchange(slab, group).wait()


def median3d_clahe_slab(slab: torch.Tensor, clip_limit: float = 2.0, grid_size: tuple = (8, 8), *,
                        mode: str = "nearest", value_range=None, group=None) -> torch.Tensor:
    """BASELINE.json config 3 on this rank's z-slab: 3x3x3 median (skimage.filters.median semantics)
    followed by per-slice CLAHE.  The halo exchange overlaps the median of the interior planes; the
    result is bit-identical to processing the unsharded volume on one GPU."""
    from .enhance import equalize_clahe
    from .filters import median

    if slab.dim() != 3:
        raise ValueError("expected a (D, H, W) slab")
    d = slab.shape[0]
    slab = slab.contiguous()
    med = torch.empty_like(slab)
    if d == 0:
        return med
    ex = start_z_halo_exchange(slab, group)
    if d >= 3:  # interior planes need no remote data
        median(slab[1:d - 1], out=med[1:d - 1], mode=mode, halo_lo=slab[0], halo_hi=slab[d - 1])
    halo_lo, halo_hi = ex.wait()
    if d == 1:
        median(slab, out=med, mode=mode, halo_lo=halo_lo, halo_hi=halo_hi)
    else:
        median(slab[0:1], out=med[0:1], mode=mode, halo_lo=halo_lo, halo_hi=slab[1])
        median(slab[d - 1:d], out=med[d - 1:d], mode=mode, halo_lo=slab[d - 2], halo_hi=halo_hi)
    return equalize_clahe(med.unsqueeze(1), clip_limit, grid_size, value_range=value_range).squeeze(1)
