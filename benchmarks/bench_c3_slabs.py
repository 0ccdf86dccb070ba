#!/usr/bin/env python
"""BASELINE.json config 3 on N GPUs: 512^3 int16 volume, 3x3x3 median + per-slice CLAHE, z-slab sharded with
one NCCL halo plane per interior face (strong scaling: the volume is fixed, each rank owns 512/N planes).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        benchmarks/bench_c3_slabs.py [--d 512] [--steps 20]
Rank 0 prints one JSON line: Mvoxel/s of the whole volume (CUDA events, max over ranks) and whether the
concatenated slabs equal the unsharded single-GPU result bit for bit (checked through a 64-bit checksum of
every rank's slab against the same planes computed on rank 0 from the full volume when --verify is given)."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mie_b200 as M  # noqa: E402
from mie_b200 import synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--d", type=int, default=512)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--verify", action="store_true")
    ap.add_argument("--eager", action="store_true", help="time eager calls instead of the captured SlabPlan")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    vol = synthetic.phantom_volume((args.d, 512, 512), np.int16, seed=0)
    z0, z1 = M.shard_range(args.d, world, rank)
    slab = torch.from_numpy(vol[z0:z1].copy()).to(dev)
    vr = None   # integer <-> [0,1] mapping = the full dtype range (an HU window is an option)

    plan = None if args.eager else M.SlabPlan(slab, 2.0, (8, 8), value_range=vr)

    def step():
        if plan is not None:
            return plan.replay()
        return M.median3d_clahe_slab(slab, 2.0, (8, 8), value_range=vr)

    for _ in range(3):
        out = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    same = None
    if args.verify:
        ok = torch.ones(1, device=dev)
        full = torch.from_numpy(vol).to(dev)
        # unsharded reference computed on this rank's own GPU (no halos: the whole volume is local)
        from mie_b200 import enhance, filters

        med = filters.median(full)
        ref = enhance.equalize_clahe(med.unsqueeze(1), 2.0, (8, 8), value_range=vr).squeeze(1)
        ok[0] = float(torch.equal(ref[z0:z1], out))
        if world > 1:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        same = bool(ok.item() == 1.0)
    if rank == 0:
        vox = args.d * 512 * 512
        t = float(ms.item())
        print(json.dumps({"config": f"C3 median3d 3x3x3 + per-slice CLAHE, {args.d}x512x512 i16, {world} z-slab(s)",
                          "n_gpus": world, "ms": round(t, 4), "mvoxel_s": round(vox / t / 1e3, 1), "scaling": "strong", "launch": "eager" if args.eager else "CUDA graph (SlabPlan)",
                          "halo_bytes_per_face": 512 * 512 * 2, "bit_identical_to_unsharded": same}), flush=True)
    if plan is not None:
        plan.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
