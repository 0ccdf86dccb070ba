#!/usr/bin/env python
"""e2e of the host pipeline at N ranks (torchrun), taper on / off, against the copy-only ceiling on the same buffers:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P benchmarks/e2e_n.py"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mie_b200 import synthetic  # noqa: E402
from mie_b200.loader import HostSlicePipeline  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
d = None
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
    d = dist
x = torch.from_numpy(synthetic.phantom((256, 1, 512, 512), np.uint16, seed=rank)).pin_memory()
y = torch.empty_like(x).pin_memory()
out = {"n": world}
for rep in range(2):
    for taper in (False, True):
        for chunk in (32, 64):
            pipe = HostSlicePipeline(dev, (512, 512), torch.uint16, chunk=chunk, taper=taper)
            for _ in range(3):
                pipe.run(x, y)
            torch.cuda.synchronize()
            if d: d.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                pipe.run(x, y)
            e1.record(); torch.cuda.synchronize()
            out[f"rep{rep}_chunk{chunk}_taper{int(taper)}_ms"] = round(bench._max_over_ranks(d, dev, e0.elapsed_time(e1) / 5), 3)
    out[f"rep{rep}_ceiling"] = bench.copy_ceiling(dev, x, y, d)
if rank == 0:
    print(json.dumps(out))
if d:
    d.barrier(); d.destroy_process_group()
