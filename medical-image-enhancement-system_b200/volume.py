"""Sharding of the enhancement path across the GPUs of one box.

  * slice batches (BASELINE.json configs 2, 4): a contiguous split of the batch
    dimension — `shard_range` — and no communication at all;
  * volumes (config 3): z-slabs.  The 3x3x3 median needs ONE neighbouring plane per
    interior slab face.  On GPUs the exchange is the C ABI's mie_halo_exchange_z
    (include/mie.h): raw ncclSend / ncclRecv in one group on a side stream, issued on
    the process group's own communicator (ProcessGroupNCCL._comm_ptr()), so the data
    plane is native code and graph-capturable; with a non-NCCL backend (gloo in the CPU
    tests) the same planes go through torch.distributed point-to-point operations.
    `median3d_clahe_slab` overlaps the exchange with the median of the slab's interior
    planes.  CLAHE is per slice and needs nothing further.

One process per GPU (torchrun); every function here is a no-op wrapper when
torch.distributed is not initialised (world size 1).
"""
from __future__ import annotations

import weakref

import torch
import torch.distributed as dist

from ._ffi import check, lib

__all__ = ["shard_range", "exchange_z_halos", "start_z_halo_exchange", "median3d_clahe_slab", "SlabPlan", "PeerSlabPlan",
           "map_peer_halos", "PeerPlane", "world_info", "nccl_comm_ptr", "close_all_plans"]


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
    """Handle of an in-flight halo exchange; .wait() returns (halo_lo, halo_hi).  For the native NCCL path
    `event` marks the end of the exchange on its side stream and wait() makes the CURRENT stream wait for it
    (stream-ordered, no host blocking); for the torch.distributed path wait() waits on the requests."""

    def __init__(self, reqs, lo, hi, event=None, device=None):
        self._reqs, self._lo, self._hi, self._event, self._device = reqs, lo, hi, event, device

    def wait(self):
        for r in self._reqs:
            r.wait()
        self._reqs = []
        if self._event is not None:
            torch.cuda.current_stream(self._device).wait_event(self._event)
            self._event = None
        return self._lo, self._hi


_side_streams = {}


def _side_stream(device: torch.device) -> torch.cuda.Stream:
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=device)
    return _side_streams[key]


def nccl_comm_ptr(device: torch.device, group=None) -> int:
    """ncclComm_t of the process group's NCCL backend for `device` as an integer (0 when the backend is not NCCL,
    this torch build does not expose it, or the native exchange is unavailable).  The communicator is created
    lazily by torch: if it does not exist yet one tiny all_reduce creates it."""
    if not (dist.is_available() and dist.is_initialized()) or device.type != "cuda":
        return 0
    if dist.get_backend(group) != "nccl" or not lib().mie_halo_exchange_available():
        return 0
    pg = group if group is not None else dist.distributed_c10d._get_default_group()
    try:
        backend = pg._get_backend(device)
        get = backend._comm_ptr
    except (AttributeError, RuntimeError):
        return 0
    for attempt in range(2):
        try:
            ptr = int(get())
        except RuntimeError:
            ptr = 0
        if ptr:
            return ptr
        if attempt == 0:
            dist.all_reduce(torch.zeros(1, device=device), group=group)   # creates the communicator
            torch.cuda.synchronize(device)
    return 0


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
    comm = nccl_comm_ptr(slab.device, group) if slab.is_cuda else 0
    if comm:
        # native data plane: mie_halo_exchange_z on a side stream, ordered after the slab's producer
        cur = torch.cuda.current_stream(slab.device)
        side = _side_stream(slab.device)
        side.wait_stream(cur)
        with torch.cuda.device(slab.device), torch.cuda.stream(side):
            check(lib().mie_halo_exchange_z(
                comm, rank, world, first.data_ptr(), last.data_ptr(), lo.data_ptr() if lo is not None else None,
                hi.data_ptr() if hi is not None else None, h * w * slab.element_size(), side.cuda_stream))
            ev = torch.cuda.Event()
            ev.record(side)
        if not torch.cuda.is_current_stream_capturing():   # a capture's private pool outlives both streams anyway
            for t in (first, last, lo, hi):
                if t is not None:
                    t.record_stream(side)
        ex = _HaloExchange([], lo, hi, ev, slab.device)
        ex._keep = (first, last)
        return ex
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
        self.graph = None
        _live_plans.add(self)
        _guard_process_group_teardown()
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
        if self.graph is None:
            raise RuntimeError("SlabPlan is closed")
        self.graph.replay()
        return self.out

    def close(self) -> None:
        """Release the captured graph.  A live graph that holds captured NCCL operations makes the process-group
        teardown hang, so plans close themselves: on leaving a `with SlabPlan(...) as plan:` block, on garbage
        collection, and — for plans still alive then — when torch.distributed.destroy_process_group() is called
        (the first SlabPlan wraps that function once; see _guard_process_group_teardown)."""
        if self.graph is not None:
            torch.cuda.synchronize(self.slab.device)
            self.graph.reset()
            self.graph = None
            self.out = None
        _live_plans.discard(self)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PeerPlane:
    """One (H, W) plane of a neighbour rank's slab, mapped into this process by CUDA IPC: a raw device pointer with the
    few tensor attributes filters.median() inspects."""

    is_cuda = True

    def __init__(self, ptr: int, shape, dtype, device):
        self._ptr, self.shape, self.dtype, self.device = int(ptr), tuple(shape), dtype, device

    def data_ptr(self) -> int:
        return self._ptr

    def is_contiguous(self) -> bool:
        return True


class _PeerMapping:
    def __init__(self, base: int):
        self.base = base

    def close(self):
        if self.base:
            lib().mie_ipc_close(self.base)
            self.base = 0


def map_peer_halos(slab: torch.Tensor, group=None):
    """Map the neighbour ranks' slabs into this process (CUDA IPC: mie_ipc_export / mie_ipc_open, include/mie.h) and
    return (halo_lo, halo_hi, keep): the lower neighbour's LAST plane and the upper neighbour's FIRST plane as PeerPlane
    objects — memory of the neighbours' GPUs, readable by this GPU's kernels over NVLink — and the mappings to close.
    None at a volume face.  Collective: every rank of the group calls it with its contiguous (D, H, W) slab (D >= 1);
    one box only.  The caller keeps `slab` alive and in place while a peer may read it."""
    import ctypes as C

    rank, world = world_info(group)
    if world == 1:
        return None, None, ()
    if not slab.is_cuda or slab.dim() != 3 or not slab.is_contiguous() or slab.shape[0] < 1:
        raise ValueError("map_peer_halos needs a contiguous, non-empty (D, H, W) CUDA slab")
    def all_ok(ok: bool, what: str):
        """Every rank learns whether EVERY rank succeeded, so that a failure is raised by all of them together (a rank that
        raised alone would leave the others waiting in the next collective)."""
        flag = torch.tensor([1.0 if ok else 0.0], device=slab.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if flag.item() != 1.0:
            raise RuntimeError(f"peer halos unavailable: {what} failed on at least one rank")

    handle = (C.c_ubyte * 64)()
    off = C.c_int64(0)
    with torch.cuda.device(slab.device):
        rc = lib().mie_ipc_export(slab.data_ptr(), handle, C.byref(off))
    all_ok(rc == 0, "exporting the slab's CUDA IPC handle")
    mine = (bytes(handle), int(off.value), tuple(slab.shape), slab.device.index)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine, group=group)
    lo = hi = None
    keep = []
    plane_bytes = slab.shape[1] * slab.shape[2] * slab.element_size()
    ok = True
    with torch.cuda.device(slab.device):
        for nb in (rank - 1, rank + 1):
            if 0 <= nb < world and ok:
                hbytes, offset, shape, dev_index = gathered[nb]
                base = C.c_void_p(0)
                buf = (C.c_ubyte * 64).from_buffer_copy(hbytes)
                if tuple(shape[1:]) != tuple(slab.shape[1:]) or lib().mie_ipc_open(buf, C.byref(base)) != 0:
                    ok = False
                    break
                keep.append(_PeerMapping(base.value))
                first = base.value + offset
                if nb < rank:
                    lo = PeerPlane(first + (shape[0] - 1) * plane_bytes, slab.shape[1:], slab.dtype, torch.device("cuda", dev_index))
                else:
                    hi = PeerPlane(first, slab.shape[1:], slab.dtype, torch.device("cuda", dev_index))
    try:
        all_ok(ok, "mapping a neighbour's slab (no peer path, or plane shapes differ)")
    except RuntimeError:
        for m in keep:
            m.close()
        raise
    return lo, hi, tuple(keep)                    # (the all-reduce above is also the "everybody has mapped" barrier)


class PeerSlabPlan:
    """BASELINE.json config 3 on this rank's z-slab with the halos READ IN PLACE: the two neighbour slabs are mapped into
    this process once (map_peer_halos), and the 3x3x3 median takes their boundary planes as its halo pointers — peer loads
    over NVLink / NVSwitch inside the kernel.  No exchange launch, no split into interior and boundary planes: the step is
    ONE median launch plus the CLAHE launches, captured into a CUDA graph.

        plan = PeerSlabPlan(slab)        # collective: maps the neighbours, captures the step
        slab.copy_(next_part); dist.barrier()   # the NEIGHBOURS' slabs must be complete before a replay reads them
        out = plan.replay()

    Synchronisation is the caller's: a replay reads the neighbours' INPUT slabs, so every rank must have finished
    refilling its slab (and must not refill it again) while a neighbour's replay is in flight.  Results are bit-identical
    to SlabPlan / the unsharded volume."""

    def __init__(self, slab: torch.Tensor, clip_limit: float = 2.0, grid_size: tuple = (8, 8), *,
                 mode: str = "nearest", value_range=None, group=None):
        from .enhance import equalize_clahe
        from .filters import median

        if not slab.is_cuda or slab.dim() != 3 or not slab.is_contiguous() or slab.shape[0] < 1:
            raise ValueError("PeerSlabPlan needs a contiguous, non-empty (D, H, W) CUDA slab")
        self.slab = slab
        self.graph = None
        self.halo_lo, self.halo_hi, self._keep = map_peer_halos(slab, group)
        self._group = group

        def step():
            med = median(slab, mode=mode, halo_lo=self.halo_lo, halo_hi=self.halo_hi, peer_halos=True)
            return equalize_clahe(med.unsqueeze(1), clip_limit, grid_size, value_range=value_range).squeeze(1)

        with torch.cuda.device(slab.device):
            step()                                                   # lazy initialisations outside the capture
            torch.cuda.synchronize(slab.device)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.out = step()
            torch.cuda.synchronize(slab.device)

    def replay(self) -> torch.Tensor:
        if self.graph is None:
            raise RuntimeError("PeerSlabPlan is closed")
        self.graph.replay()
        return self.out

    def close(self) -> None:
        """Collective: every rank finishes its replays, then the mappings of the neighbours' slabs are dropped."""
        if self.graph is not None:
            torch.cuda.synchronize(self.slab.device)
            self.graph.reset()
            self.graph = None
            self.out = None
            if dist.is_available() and dist.is_initialized():
                dist.barrier(group=self._group)                      # no neighbour is still reading this rank's slab
            self.halo_lo = self.halo_hi = None
            for m in self._keep:
                m.close()
            self._keep = ()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


_live_plans = weakref.WeakSet()
_teardown_guarded = False


def close_all_plans() -> None:
    """Close every live SlabPlan of this process (idempotent)."""
    for p in list(_live_plans):
        p.close()


def _guard_process_group_teardown() -> None:
    """Wrap torch.distributed.destroy_process_group ONCE so that it first closes the live SlabPlans: a forgotten
    close() can then no longer hang the teardown."""
    global _teardown_guarded
    if _teardown_guarded:
        return
    _teardown_guarded = True
    original = dist.destroy_process_group

    def destroy_process_group(*args, **kwargs):
        close_all_plans()
        return original(*args, **kwargs)

    destroy_process_group.__wrapped__ = original
    dist.destroy_process_group = destroy_process_group
    try:
        dist.distributed_c10d.destroy_process_group = destroy_process_group
    except AttributeError:
        pass
