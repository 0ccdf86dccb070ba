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

__all__ = ["shard_range", "exchange_z_halos", "start_z_halo_exchange", "median3d_clahe_slab", "SlabPlan", "world_info"]


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


class SlabPlan:
    """median3d_clahe_slab on a FIXED slab buffer, captured once into a CUDA graph — the halo send / recv
    (NCCL supports stream capture), the three median launches, the CLAHE launches and the allocations they
    make — and replayed per volume: at 8 GPUs a 64-plane slab needs ~0.1 ms of kernels, and a dozen eager
    launches plus the Python around them cost several times that.

        plan = SlabPlan(slab)            # warm-up call (creates the NCCL pair communicators) + capture
        slab.copy_(next_volume_part)     # refill the same buffer
        out = plan.replay()              # this rank's part of the result (owned by the plan)

    Every rank of the group must build the plan and replay it the same number of times."""

    def __init__(self, slab: torch.Tensor, clip_limit: float = 2.0, grid_size: tuple = (8, 8), *,
                 mode: str = "nearest", value_range=None, group=None):
        if not slab.is_cuda or slab.dim() != 3 or not slab.is_contiguous():
            raise ValueError("SlabPlan needs a contiguous (D, H, W) CUDA slab")
        self.slab = slab
        self._args = (clip_limit, grid_size)
        self._kw = dict(mode=mode, value_range=value_range, group=group)
        with torch.cuda.device(slab.device):
            median3d_clahe_slab(slab, *self._args, **self._kw)       # eager: NCCL communicators, lazy inits
            torch.cuda.synchronize(slab.device)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.out = median3d_clahe_slab(slab, *self._args, **self._kw)
            torch.cuda.synchronize(slab.device)

    def replay(self) -> torch.Tensor:
        self.graph.replay()
        return self.out

    def close(self) -> None:
        """Release the captured graph.  Call it (on every rank) before destroy_process_group(): a live graph
        that holds captured NCCL operations makes the process-group teardown hang."""
        if self.graph is not None:
            torch.cuda.synchronize(self.slab.device)
            self.graph.reset()
            self.graph = None
            self.out = None
