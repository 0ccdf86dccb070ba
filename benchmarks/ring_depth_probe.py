#!/usr/bin/env python
"""Throughput of the config-2 chain with 1..6 batches in flight (ChainRing depth): python benchmarks/ring_depth_probe.py"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import mie_b200 as M
from mie_b200 import synthetic
dev = torch.device("cuda:0")
for depth in (1, 2, 3, 4, 6):
    xs = [torch.from_numpy(synthetic.phantom((256, 1, 512, 512), np.uint16, seed=s)).to(dev) for s in range(depth)]
    ring = M.ChainRing(xs, M.ChainConfig())
    def run(k):
        ring.begin()
        for i in range(k):
            ring.replay(i)
        ring.join()
    run(12); torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(60); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 60)
    print(f"depth {depth}: {best:.4f} ms per 256-slice batch", flush=True)
    del ring, xs
    torch.cuda.empty_cache()
