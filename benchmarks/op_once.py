#!/usr/bin/env python
"""Runs ONE operator a few times on its BASELINE.json config shape (for ncu captures and quick timings):
    python benchmarks/op_once.py median3d|median2d|median2d_f32|median5|mse|ssim|equalize|clahe|clahe16|gauss|unsharp|bilateral|nlm [reps]
Prints one JSON line with the CUDA-event time per call (eager calls; never quote a number taken under ncu)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mie_b200 as M  # noqa: E402
from mie_b200 import synthetic  # noqa: E402

op = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda:0")


def batch2():
    return torch.from_numpy(synthetic.phantom((256, 1, 512, 512), np.uint16, 0)).to(dev)


if op == "median3d":
    d = int(os.environ.get("MIE_D", "256"))
    v = torch.from_numpy(synthetic.phantom_volume((d, 512, 512), np.int16, 0)).to(dev)
    fn, px = (lambda: M.median(v)), d * 512 * 512
elif op == "median2d":
    x = batch2(); fn, px = (lambda: M.median_blur(x, 3)), x.numel()
elif op == "median2d_f32":
    x = batch2().to(torch.float32) / 65535.0; fn, px = (lambda: M.median_blur(x, 3)), x.numel()
elif op == "median5":
    x = batch2(); fn, px = (lambda: M.median_blur(x, 5)), x.numel()
elif op in ("mse", "ssim"):
    x = batch2(); y = M.unsharp_mask(x, 9, 1.0)
    fn, px = ((lambda: M.mse(x, y)) if op == "mse" else (lambda: M.ssim(x, y))), x.numel()
elif op == "equalize":
    x = batch2(); fn, px = (lambda: M.equalize(x)), x.numel()
elif op == "clahe":
    x = batch2(); fn, px = (lambda: M.equalize_clahe(x, 2.0, (8, 8))), x.numel()
elif op == "clahe16":
    x = batch2()[:64]; fn, px = (lambda: M.equalize_clahe(x, 2.0, (8, 8), semantics="opencv")), x.numel()
elif op == "gauss":
    x = batch2(); fn, px = (lambda: M.gaussian_blur2d(x, 9, 1.0)), x.numel()
elif op == "unsharp":
    x = batch2(); fn, px = (lambda: M.unsharp_mask(x, 9, 1.0)), x.numel()
elif op == "bilateral":
    x = torch.from_numpy(synthetic.phantom((2, 1, 4096, 4096), np.uint16, 0)).to(dev)
    fn, px = (lambda: M.bilateral_blur(x, 9, 0.1, (1.5, 1.5), out_dtype=torch.float32)), x.numel()
elif op == "nlm":
    x = torch.from_numpy(synthetic.phantom((64, 1, 256, 256), np.uint16, 0)).to(dev)
    fn, px = (lambda: M.denoise_nl_means(x, 7, 11, 0.1)), x.numel()
else:
    raise SystemExit("unknown op " + op)

for _ in range(2):
    fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    fn()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(json.dumps({"op": op, "ms": round(ms, 4), "mpixel_s": round(px / ms / 1e3, 1), "pixels": px}))
