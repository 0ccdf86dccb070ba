#!/usr/bin/env python
"""scikit-image compatibility operators (float64 arithmetic in upstream's order) on config-sized inputs: CUDA-event timing.
    python benchmarks/sk_quick.py"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mie_b200 as M  # noqa: E402
from mie_b200 import skimage_compat as S, synthetic  # noqa: E402

dev = torch.device("cuda:0")
x = torch.from_numpy(synthetic.phantom((64, 1, 512, 512), np.uint16, 0)).to(dev)


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


OPS = {
    "sk.equalize_adapthist 64x512x512 (one call, float64 out)": (lambda: S.equalize_adapthist(x), 64 * 512 * 512),
    "sk.equalize_adapthist 1x512x512 (latency)": (lambda: S.equalize_adapthist(x[0, 0]), 512 * 512),
    "sk.equalize_hist 64x512x512 (one call, float64 out)": (lambda: S.equalize_hist(x), 64 * 512 * 512),
    "sk.equalize_hist 1x512x512 (latency)": (lambda: S.equalize_hist(x[0, 0]), 512 * 512),
    "sk.denoise_bilateral 1x512x512 (win 7, 10000-bin colour LUT)": (lambda: S.denoise_bilateral(x[0, 0], sigma_color=0.05, sigma_spatial=1), 512 * 512),
    "sk.gaussian 64x512x512 sigma=1": (lambda: S.gaussian(x, 1.0), 64 * 512 * 512),
    "sk.unsharp_mask 64x512x512 radius=1 amount=1": (lambda: S.unsharp_mask(x, 1.0, 1.0), 64 * 512 * 512),
    "denoise_nl_means fast 16x256x256": (lambda: M.denoise_nl_means(x[:16, :, :256, :256].contiguous(), 7, 11, 0.1), 16 * 256 * 256),
    "denoise_nl_means slow 16x256x256": (lambda: M.denoise_nl_means(x[:16, :, :256, :256].contiguous(), 7, 11, 0.1, fast_mode=False), 16 * 256 * 256),
}
for name, (fn, px) in OPS.items():
    try:
        ms = timed(fn, 3)
        print(json.dumps({"op": name, "ms": round(ms, 3), "mpixel_s": round(px / ms / 1e3, 1)}), flush=True)
    except Exception as exc:   # an operator signature this script does not know: report, keep going
        print(json.dumps({"op": name, "error": f"{type(exc).__name__}: {exc}"[:200]}), flush=True)
