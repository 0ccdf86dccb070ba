"""Small calls of this session's kernels for compute-sanitizer (memcheck, racecheck, synccheck): cluster equalisation
(1 / 2 / 4 / 8 CTAs, ragged rows), lean bilateral (all windows, ragged tiles, every border), fused bilateral -> CLAHE,
slow-mode NLM, bounded 65 536-bin CLAHE LUTs (both halves live / upper half retired)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import mie_b200 as M
from mie_b200 import synthetic
dev = torch.device("cuda:0")
for shape, dt in [((3, 1, 64, 128), np.uint16), ((2, 1, 200, 256), np.int16), ((2, 1, 512, 512), np.uint16),
                  ((1, 1, 100, 64), np.uint8), ((1, 1, 1024, 1024), np.uint16)]:
    x = torch.from_numpy(synthetic.make("P", shape, dt, 1)).to(dev)
    a = M.equalize(x)
    with M.kernel_policy("equalize_three_pass"):
        b = M.equalize(x)
    assert torch.equal(a.view(torch.uint8), b.view(torch.uint8)), shape
    M.equalize(x, out_dtype=torch.float32)
x = torch.from_numpy(synthetic.phantom((2, 1, 70, 100), np.uint16, 1)).to(dev)
for k in (3, 5, 7, 9):
    for border in ("reflect", "replicate", "constant", "circular"):
        M.bilateral_blur(x, k, 0.1, (1.5, 1.5), border)
xb = torch.from_numpy(synthetic.phantom((2, 1, 128, 256), np.uint16, 1)).to(dev)
M.bilateral_clahe(xb, 9, 0.1, (1.5, 1.5), 2.0, (2, 4))
M.bilateral_clahe(xb, 5, 0.1, (1.5, 1.5), 2.0, (4, 8), "replicate", out_dtype=torch.float32)
M.denoise_nl_means(x[:1, :, :40, :52], 5, 4, 0.1, fast_mode=False, out_dtype=torch.float64)
M.denoise_nl_means(x[:1, :, :33, :21].contiguous(), 7, 3, 0.1, fast_mode=False)
rng = np.random.default_rng(0)
for top in (4095, 40000, 65535, 0):
    y = torch.from_numpy(rng.integers(0, top + 1, (20, 1, 128, 128), dtype=np.uint16)).to(dev)
    a = M.equalize_clahe(y, 2.0, (4, 4), semantics="opencv")
    with M.kernel_policy("clahe16_full_luts"):
        b = M.equalize_clahe(y, 2.0, (4, 4), semantics="opencv")
    assert torch.equal(a.view(torch.int16), b.view(torch.int16)), top
M.equalize_clahe(x, 2.0, (2, 2))
torch.cuda.synchronize()
print("sanitize run ok")
