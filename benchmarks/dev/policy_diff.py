"""Dev probe: tuned vs generic vs oracle for every op with a policy bit (prints mismatch counts)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import mie_b200 as M
import oracle as O
from mie_b200 import synthetic
dev = torch.device("cuda:0")
for dtype in (np.uint16, np.int16, np.uint8):
    xn = synthetic.make("P", (6, 1, 256, 512), dtype, 11)
    x = torch.from_numpy(xn).to(dev)
    x01 = O.to01(xn)
    ops = [
        ("generic_gauss", "gauss", lambda: M.gaussian_blur2d(x, 9, 1.0), lambda: O.from01(O.gaussian_blur2d(x01, 9, 1.0), dtype)),
        ("generic_gauss", "unsharp", lambda: M.unsharp_mask(x, 9, 1.0), lambda: O.from01(O.unsharp_mask(x01, 9, 1.0), dtype)),
        ("generic_clahe", "clahe", lambda: M.equalize_clahe(x, 2.0, (4, 8)), lambda: O.from01(O.equalize_clahe(x01, 2.0, (4, 8)), dtype)),
        ("clahe_float_rules", "clahe", lambda: M.equalize_clahe(x, 2.0, (4, 8)), None),
        ("generic_equalize", "equalize", lambda: M.equalize(x), lambda: O.from01(O.equalize(x01), dtype)),
        ("equalize_float_rules", "equalize", lambda: M.equalize(x), None),
        ("generic_median", "median3", lambda: M.median_blur(x, 3), lambda: O.median_blur(xn, 3)),
        ("generic_median", "median5", lambda: M.median_blur(x, 5), lambda: O.median_blur(xn, 5)),
        ("generic_bilateral", "bilateral", lambda: M.bilateral_blur(x[:2], 5, 0.1, (1.5, 1.5)), None),
    ]
    for pol, name, fn, ref in ops:
        t = fn().cpu().numpy()
        with M.kernel_policy(pol):
            g = fn().cpu().numpy()
        r = ref() if ref else None
        msg = f"{np.dtype(dtype).name:7s} {name:10s} policy={pol:22s} tuned!=generic: {int((t != g).sum())}"
        if r is not None:
            msg += f"  tuned!=oracle: {int((t != r.reshape(t.shape)).sum())}  generic!=oracle: {int((g != r.reshape(g.shape)).sum())}"
        if (t != g).any():
            idx = np.argwhere(t != g)[:4]
            msg += f"  first diffs at {idx.tolist()}"
        print(msg, flush=True)
