#!/usr/bin/env python
"""Secondary benchmark: every BASELINE.json config on ONE GPU, resident data, CUDA events.

bench.py reports only the headline (config 2); this script records the other configs (which are
parity-test cases, not bench lines) so that DESIGN.md / profiles/ can state what bounds each of them.
    python benchmarks/bench_configs.py [--quick] > gpurun_out/configs.jsonl
"""
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


def timed(fn, reps, warm=3):
    """ms per call, CUDA events.  The call (kernels + the output allocation it makes) is captured once into
    a CUDA graph and replayed, so that sub-0.1 ms operators are not timed through Python / allocator
    overhead; operators that cannot be captured are timed eagerly."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    run = fn
    try:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        run = g.replay
        run()
    except Exception:
        run = fn
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def report(name, pixels, ms, bytes_per_px=4, **extra):
    gbs = pixels * bytes_per_px / (ms * 1e-3) / 1e9
    line = {"config": name, "ms": round(ms, 4), "mpixel_s": round(pixels / (ms * 1e-3) / 1e6, 1),
            "alg_GBps": round(gbs, 1), "frac_of_measured_6548.8": round(gbs / 6548.8, 4), **extra}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    reps = 5 if args.quick else 20

    # C1: single 512x512 uint16 slice, CLAHE 8x8 clip 2.0 (latency-bound: report microseconds)
    x1 = torch.from_numpy(synthetic.phantom((1, 1, 512, 512), np.uint16, 0)).to(dev)
    ms = timed(lambda: M.equalize_clahe(x1, 2.0, (8, 8)), 200)
    report("C1 clahe 1x512x512 u16 (python call -> 2 launches)", 512 * 512, ms, latency_us=round(ms * 1e3, 1))
    ms = timed(lambda: M.enhance_chain(x1), 200)
    report("C1' chain 1x512x512 u16 (3 launches)", 512 * 512, ms, latency_us=round(ms * 1e3, 1))
    ms = timed(lambda: M.equalize_clahe(x1, 2.0, (8, 8), semantics="opencv"), 100)
    report("C1 clahe 1x512x512 u16, OpenCV semantics, 65536 bins (cv2: ~20 ms on 8 threads)", 512 * 512, ms,
           latency_us=round(ms * 1e3, 1))

    # C2 pieces: the standalone ops on the config-2 batch
    x2 = torch.from_numpy(synthetic.phantom((256, 1, 512, 512), np.uint16, 0)).to(dev)
    px2 = 256 * 512 * 512
    report("C2 chain fused 256x512x512 u16", px2, timed(lambda: M.enhance_chain(x2), reps))
    report("C2 gaussian_blur2d K=9 u16->u16", px2, timed(lambda: M.gaussian_blur2d(x2, 9, 1.0), reps))
    report("C2 unsharp_mask K=9 u16->u16", px2, timed(lambda: M.unsharp_mask(x2, 9, 1.0), reps))
    report("C2 equalize_clahe 8x8 u16->u16", px2, timed(lambda: M.equalize_clahe(x2, 2.0, (8, 8)), reps))
    report("C2 equalize (global) u16->u16", px2, timed(lambda: M.equalize(x2), reps))
    report("C2 median_blur 3x3 u16", px2, timed(lambda: M.median_blur(x2, 3), reps))
    report("C2 equalize_clahe 8x8 u16, OpenCV semantics, 65536 bins (LUTs bounded by the batch maximum, 512 MB LUT workspace)", px2,
           timed(lambda: M.equalize_clahe(x2, 2.0, (8, 8), semantics="opencv"), max(reps // 4, 2)))
    del x2

    # C5: non-local means 7x7 patches, search radius 11, on 256x256 slices (batch 512; sample when --quick)
    nb5 = 64 if args.quick else 512
    x5 = torch.from_numpy(synthetic.phantom((nb5, 1, 256, 256), np.uint16, 0)).to(dev)
    report(f"C5 denoise_nl_means 7x7 d=11 {nb5}x256x256 u16", nb5 * 256 * 256,
           timed(lambda: M.denoise_nl_means(x5, 7, 11, 0.1), 2, warm=1))
    del x5

    # C3: 512^3 int16 volume, 3x3x3 median + per-slice CLAHE (one GPU = one slab with no halos)
    d = 128 if args.quick else 512
    v = torch.from_numpy(synthetic.phantom_volume((d, 512, 512), np.int16, 0)).to(dev)
    report(f"C3 median3d 3x3x3 {d}x512x512 i16", d * 512 * 512, timed(lambda: M.median(v), max(reps // 4, 2)))
    report(f"C3 median3d + per-slice CLAHE {d}x512x512 i16", d * 512 * 512,
           timed(lambda: M.median3d_clahe_slab(v, 2.0, (8, 8)), max(reps // 4, 2)))
    del v

    # C4: 4096x4096 uint16 radiographs, 9x9 bilateral + CLAHE 16x16 (sample of the 64-image batch)
    nb = 2 if args.quick else 8
    x4 = torch.from_numpy(synthetic.phantom((nb, 1, 4096, 4096), np.uint16, 0)).to(dev)
    px4 = nb * 4096 * 4096

    def c4():
        b = M.bilateral_blur(x4, 9, 0.1, (1.5, 1.5), out_dtype=torch.float32)
        return M.equalize_clahe(b, 2.0, (16, 16))

    report(f"C4 bilateral 9x9 {nb}x4096x4096 u16->f32", px4,
           timed(lambda: M.bilateral_blur(x4, 9, 0.1, (1.5, 1.5), out_dtype=torch.float32), 2, warm=1), bytes_per_px=6)
    report(f"C4 bilateral + CLAHE 16x16 {nb}x4096x4096 (sample of 64)", px4, timed(c4, 2, warm=1), bytes_per_px=6)


if __name__ == "__main__":
    main()
