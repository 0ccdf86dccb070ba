#!/usr/bin/env python
"""Approximate-exponential (MUFU.EX2) bilateral against the exact-polynomial kernel: error statistics of the filter
output and of the fused bilateral -> CLAHE chain (config 4), and timings on an 8-image sample of the config-4 batch.
    python benchmarks/bilateral_approx_probe.py [n_images]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mie_b200 as M  # noqa: E402
from mie_b200 import synthetic  # noqa: E402

dev = torch.device("cuda:0")
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 8
x = torch.from_numpy(synthetic.phantom((nb, 1, 4096, 4096), np.uint16, 0)).to(dev)


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for kind in ("P", "U"):
    xs = torch.from_numpy(synthetic.make(kind, (2, 1, 1024, 1024), np.uint16, 3)).to(dev)
    a = M.bilateral_blur(xs, 9, 0.1, (1.5, 1.5), out_dtype=torch.float32)
    with M.kernel_policy("bilateral_exact_exp"):
        e = M.bilateral_blur(xs, 9, 0.1, (1.5, 1.5), out_dtype=torch.float32)
    d = (a - e).abs()
    rel = (d / e.abs().clamp_min(1e-3)).max().item()
    qa = M.bilateral_blur(xs, 9, 0.1, (1.5, 1.5)).to(torch.int32)
    with M.kernel_policy("bilateral_exact_exp"):
        qe = M.bilateral_blur(xs, 9, 0.1, (1.5, 1.5)).to(torch.int32)
    print(json.dumps({"data": kind, "max_abs": d.max().item(), "max_rel": rel, "u16_max_lsb": int((qa - qe).abs().max()),
                      "u16_frac_diff": float((qa != qe).float().mean())}))
    fa = M.bilateral_clahe(xs, 9, 0.1, (1.5, 1.5), 2.0, (4, 4)).to(torch.int32)
    with M.kernel_policy("bilateral_exact_exp"):
        fe = M.bilateral_clahe(xs, 9, 0.1, (1.5, 1.5), 2.0, (4, 4)).to(torch.int32)
    df = (fa - fe).abs()
    print(json.dumps({"data": kind, "chain_frac_diff": float((df > 0).float().mean()), "chain_max_lsb": int(df.max()),
                      "chain_frac_gt_257": float((df > 257).float().mean()), "chain_mean_lsb": float(df.float().mean())}))

t_a = timed(lambda: M.bilateral_blur(x, 9, 0.1, (1.5, 1.5), out_dtype=torch.float32))
with M.kernel_policy("bilateral_exact_exp"):
    t_e = timed(lambda: M.bilateral_blur(x, 9, 0.1, (1.5, 1.5), out_dtype=torch.float32))
c_a = timed(lambda: M.bilateral_clahe(x, 9, 0.1, (1.5, 1.5), 2.0, (16, 16)))
with M.kernel_policy("bilateral_exact_exp"):
    c_e = timed(lambda: M.bilateral_clahe(x, 9, 0.1, (1.5, 1.5), 2.0, (16, 16)))
print(json.dumps({"images": nb, "bilateral_ms_approx": round(t_a, 3), "bilateral_ms_exact": round(t_e, 3),
                  "chain_ms_approx": round(c_a, 3), "chain_ms_exact": round(c_e, 3)}))
