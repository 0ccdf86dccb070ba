#!/usr/bin/env python
"""Quick kernel-only timing of the config-2 chain (resident data, CUDA events, per-stage) with a
checksum, for kernel iteration: python benchmarks/chain_quick.py [--reps 50]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mie_b200 as M  # noqa: E402
from mie_b200 import synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=50)
    ap.add_argument("--batch", type=int, default=256)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    x = torch.from_numpy(synthetic.phantom((args.batch, 1, 512, 512), np.uint16, 0)).to(dev)
    cfg = M.ChainConfig()
    y = torch.empty_like(x)
    ws = torch.empty(M.chain_workspace_bytes(args.batch, 512, 512), dtype=torch.uint8, device=dev)

    class plan:  # noqa: N801
        @staticmethod
        def run(stages):
            M.enhance_chain(x, cfg, out=y, workspace=ws, stages=stages)
    out = {}
    import time
    inner = 20
    for name, stages in (("all", 3), ("a", 1), ("b", 2)):
        for _ in range(3):
            plan.run(stages)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(inner):
            plan.run(stages)
        out[name + "_host_us_per_call"] = round((time.perf_counter() - t0) / inner * 1e6, 1)
        torch.cuda.synchronize()
        # GPU time without host launch gaps: replay a CUDA graph holding `inner` back-to-back runs
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            with torch.cuda.graph(g, stream=side):
                for _ in range(inner):
                    plan.run(stages)
        ts = []
        for _ in range(max(args.reps // inner, 3) + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / inner)
        ts = sorted(ts[1:])
        out[name + "_ms_median"] = round(ts[len(ts) // 2], 4)
        out[name + "_ms_min"] = round(ts[0], 4)
    plan.run(3)
    torch.cuda.synchronize()
    # bit-exactness against the independent tile kernels (MIE_CHAIN_PREFER_TILES) on the same input
    y_tiles = M.enhance_chain(x, cfg, stages=3 | 8)
    out["equals_tile_kernels"] = bool(torch.equal(y.view(torch.int16), y_tiles.view(torch.int16)))
    out["differing_pixels"] = int((y.view(torch.int16) != y_tiles.view(torch.int16)).sum().item())
    out["checksum"] = int(y.view(torch.int16).to(torch.int64).sum().item() & 0xFFFFFFFF)
    px = x.numel()
    out["mpixel_s"] = round(px / out["all_ms_median"] / 1e3, 1)
    out["frac_hbm"] = round(px * 4 / (out["all_ms_median"] * 1e-3) / 1e9 / 6548.8, 4)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
