"""Host-side sharding logic on CPU (no GPU): balanced splits, and the z-halo exchange over a real
2- and 3-rank `gloo` process group, checked by stitching oracle medians of the slabs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_range_is_a_partition():
    import mie_b200 as M

    for total in (0, 1, 7, 64, 256, 512):
        for world in (1, 2, 3, 4, 8):
            spans = [M.shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        M.shard_range(10, 2, 2)


def test_halo_exchange_is_a_noop_without_a_process_group():
    import mie_b200 as M

    lo, hi = M.exchange_z_halos(torch.zeros(4, 8, 8, dtype=torch.int16))
    assert lo is None and hi is None


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, dtype_name, out_dir):
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import mie_b200 as M
    import oracle as O

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        dtype = np.dtype(dtype_name)
        rng = np.random.default_rng(42)  # same volume on every rank
        vol = rng.integers(0, 4000, (11, 12, 14)).astype(dtype)
        z0, z1 = M.shard_range(vol.shape[0], world, rank)
        slab = torch.from_numpy(vol[z0:z1].copy())
        lo, hi = M.exchange_z_halos(slab)
        assert (lo is None) == (rank == 0) and (hi is None) == (rank == world - 1)
        if lo is not None:
            assert np.array_equal(lo.numpy(), vol[z0 - 1])
        if hi is not None:
            assert np.array_equal(hi.numpy(), vol[z1])
        out = O.median3d(vol[z0:z1], "nearest", None if lo is None else lo.numpy(), None if hi is None else hi.numpy())
        np.save(os.path.join(out_dir, f"part{rank}.npy"), out)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,dtype", [(2, "int16"), (3, "uint16"), (2, "float32")])
def test_z_halo_exchange_over_gloo_reproduces_unsharded_median(tmp_path, world, dtype):
    import oracle as O

    port = _free_port()
    mp.spawn(_worker, args=(world, port, dtype, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(42)
    vol = rng.integers(0, 4000, (11, 12, 14)).astype(np.dtype(dtype))
    parts = np.concatenate([np.load(tmp_path / f"part{r}.npy") for r in range(world)])
    assert np.array_equal(parts, O.median3d(vol))


def test_peer_halo_mapping_is_a_no_op_without_a_process_group():
    """map_peer_halos (CUDA IPC mapping of the neighbour slabs) has nothing to map in a single-process run; PeerPlane is
    the duck-typed halo argument filters.median() inspects."""
    import torch

    import mie_b200 as M
    from mie_b200.volume import PeerPlane

    lo, hi, keep = M.map_peer_halos(torch.zeros((3, 4, 5), dtype=torch.int16))
    assert lo is None and hi is None and keep == ()
    p = PeerPlane(0x7F0000001000, (4, 5), torch.int16, torch.device("cuda", 3))
    assert p.data_ptr() == 0x7F0000001000 and p.shape == (4, 5) and p.is_contiguous() and p.is_cuda
    assert p.dtype == torch.int16 and p.device.index == 3
