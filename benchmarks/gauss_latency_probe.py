"""Latency of gaussian_blur2d / unsharp_mask / median_blur for small batches of 512x512 uint16 slices (graph replay).
Wrap the calls in `with M.kernel_policy('generic_gauss'):` to time the tile kernel.  python benchmarks/gauss_latency_probe.py"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import mie_b200 as M
from mie_b200 import synthetic
for nb in (1, 4, 16, 32, 64):
    x = torch.from_numpy(synthetic.phantom((nb, 1, 512, 512), np.uint16, 0)).cuda()
    for name, fn in (("gauss", lambda: M.gaussian_blur2d(x, 9, 1.0)), ("median3", lambda: M.median_blur(x, 3)),
                     ("equalize", lambda: M.equalize(x))):
        for _ in range(3): fn()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g): fn()
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(100): g.replay()
        e1.record(); torch.cuda.synchronize()
        print("batch", nb, name, round(e0.elapsed_time(e1) * 10, 1), "us")
