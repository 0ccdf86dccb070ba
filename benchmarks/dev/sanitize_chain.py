"""Small chain / CLAHE / bilateral-CLAHE / skimage calls for compute-sanitizer (memcheck, racecheck, synccheck)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import mie_b200 as M
from mie_b200 import synthetic, skimage_compat as S
dev = torch.device("cuda:0")
x = torch.from_numpy(synthetic.phantom((4, 1, 128, 256), np.uint16, 1)).to(dev)
cfg = M.ChainConfig(grid_size=(2, 4))
a = M.enhance_chain(x, cfg, stages=3 | 4)       # marching kernels
b = M.enhance_chain(x, cfg, stages=3 | 8)       # tile kernels
assert torch.equal(a.view(torch.int16), b.view(torch.int16))
M.gaussian_blur2d(x, 9, 1.0); M.unsharp_mask(x, 9, 1.0); M.equalize_clahe(x, 2.0, (2, 4)); M.equalize(x)
M.bilateral_clahe(x, 9, 0.1, (1.5, 1.5), 2.0, (2, 4))
S.equalize_adapthist(x); S.equalize_hist(x); S.denoise_bilateral(x[:1])
v = torch.from_numpy(synthetic.phantom_volume((12, 64, 64), np.int16, 0)).to(dev)
M.median(v)
torch.cuda.synchronize()
print("sanitize run ok")
