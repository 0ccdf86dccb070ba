#!/usr/bin/env python
"""Standalone operators on the config-2 batch (256 x 512x512 uint16 phantom): CUDA-graph replay timing, fraction of the
measured HBM roofline (4 B/px), and a bit-identity check of the tuned kernel against the generic one (kernel policy).
    python benchmarks/ops_quick.py [op ...]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mie_b200 as M  # noqa: E402
from mie_b200 import synthetic  # noqa: E402

PEAK = 6548.8
dev = torch.device("cuda:0")
x = torch.from_numpy(synthetic.phantom((256, 1, 512, 512), np.uint16, 0)).to(dev)
OPS = {
    "gauss": (lambda: M.gaussian_blur2d(x, 9, 1.0), "generic_gauss"),
    "unsharp": (lambda: M.unsharp_mask(x, 9, 1.0), "generic_gauss"),
    "clahe": (lambda: M.equalize_clahe(x, 2.0, (8, 8)), "generic_clahe"),
    "equalize": (lambda: M.equalize(x), "generic_equalize"),
    "median3": (lambda: M.median_blur(x, 3), "generic_median"),
    "median5": (lambda: M.median_blur(x, 5), "generic_median"),
    "clahe16": (lambda: M.equalize_clahe(x, 2.0, (8, 8), semantics="opencv"), "clahe16_full_luts"),
}


def timed(fn, reps=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for name in (sys.argv[1:] or list(OPS)):
    fn, pol = OPS[name]
    ms = timed(fn)
    tuned = fn()
    with M.kernel_policy(pol):
        generic = fn()
    same = bool(torch.equal(tuned.view(torch.int16), generic.view(torch.int16)))
    gbs = x.numel() * 4 / (ms * 1e-3) / 1e9
    print(json.dumps({"op": name, "ms": round(ms, 4), "GBps": round(gbs, 1), "frac": round(gbs / PEAK, 4),
                      "tuned_equals_generic": same}), flush=True)
