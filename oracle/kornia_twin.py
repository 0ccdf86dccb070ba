"""torch-CPU "kornia twin" — TEST INFRASTRUCTURE ONLY (never imported by the package).

kornia 0.8.2 (reference pyproject.toml:8, uv.lock:219-230) cannot be installed in
this image, so this module restates its enhancement functions with the same
*tensor-level* formulation kornia uses (pad -> unfold into tiles -> histc ->
clamp/redistribute -> cumsum -> gather -> addcmul lerps; pad -> conv2d; unfold ->
median), written from the published algorithm.  It is
structurally independent of the per-pixel restatement in mie_oracle.c, which is
the point: tests/test_oracle.py (test_kornia_clahe_oracle_matches_tensor_level_twin and
its neighbours) checks that the two agree (LUTs bit-exact, float outputs to ~1e-6),
which pins the tile/half-tile/weight bookkeeping of the oracle.  It is also the
"kornia-style torch-CPU pipeline" SURVEY.md §8(d) names as the PRIMARY CPU figure:
bench.py times it as `cpu_baseline` and as the `--impl reference` arm
(bench.py:_twin_chain).  tests/test_live_pins.py compares it with the real kornia
whenever that package is importable.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def get_gaussian_kernel1d(kernel_size: int, sigma: float) -> torch.Tensor:
    x = torch.arange(kernel_size, dtype=torch.float32) - kernel_size // 2
    if kernel_size % 2 == 0:
        x = x + 0.5
    g = torch.exp(-x.pow(2.0) / (2 * float(sigma) ** 2))
    return g / g.sum()


def _pair(v):
    return (v[0], v[1]) if isinstance(v, (tuple, list)) else (v, v)


def gaussian_blur2d(x: torch.Tensor, kernel_size, sigma, border_type: str = "reflect") -> torch.Tensor:
    """(B,C,H,W) fp32.  pad(border_type) then two grouped 1-D convolutions."""
    ky, kx = _pair(kernel_size)
    sy, sx = _pair(sigma)
    b, c, h, w = x.shape
    wx = get_gaussian_kernel1d(kx, sx).view(1, 1, 1, kx).repeat(c, 1, 1, 1)
    wy = get_gaussian_kernel1d(ky, sy).view(1, 1, ky, 1).repeat(c, 1, 1, 1)
    xp = F.pad(x, [kx // 2, kx // 2, ky // 2, ky // 2], mode=border_type)
    out = F.conv2d(xp, wx, groups=c)
    return F.conv2d(out, wy, groups=c)


def unsharp_mask(x: torch.Tensor, kernel_size, sigma, border_type: str = "reflect") -> torch.Tensor:
    return x + (x - gaussian_blur2d(x, kernel_size, sigma, border_type))


def median_blur(x: torch.Tensor, kernel_size) -> torch.Tensor:
    """zero-padded window, torch.median (lower median)."""
    ky, kx = _pair(kernel_size)
    b, c, h, w = x.shape
    xp = F.pad(x, [(kx - 1) // 2, (kx - 1) // 2, (ky - 1) // 2, (ky - 1) // 2])
    win = xp.unfold(2, ky, 1).unfold(3, kx, 1).reshape(b, c, h, w, ky * kx)
    return win.median(dim=-1)[0]


def bilateral_blur(x: torch.Tensor, kernel_size, sigma_color: float, sigma_space, border_type: str = "reflect"):
    """single-channel semantics of kornia.filters.bilateral_blur ('l1' == 'l2' for C == 1)."""
    ky, kx = _pair(kernel_size)
    sy, sx = _pair(sigma_space)
    b, c, h, w = x.shape
    xp = F.pad(x, [kx // 2, kx // 2, ky // 2, ky // 2], mode=border_type)
    win = xp.unfold(2, ky, 1).unfold(3, kx, 1).reshape(b, c, h, w, ky * kx)
    diff = win - x.unsqueeze(-1)
    color = (-0.5 / sigma_color**2 * diff.square()).exp()
    space = (get_gaussian_kernel1d(ky, sy)[:, None] * get_gaussian_kernel1d(kx, sx)[None, :]).reshape(-1)
    kern = space * color
    return (win * kern).sum(-1) / kern.sum(-1)


def equalize(x: torch.Tensor) -> torch.Tensor:
    """kornia.enhance.equalize / torchvision rule, per (B,C) plane, fp32 in [0,1]."""
    out = torch.empty_like(x)
    b, c = x.shape[:2]
    for i in range(b):
        for j in range(c):
            im = x[i, j] * 255
            histo = torch.histc(im, bins=256, min=0, max=255)
            nz = histo[histo != 0]
            step = (nz.sum() - nz[-1]) // 255 if nz.numel() else torch.tensor(0.0)
            if step == 0:
                out[i, j] = x[i, j]
                continue
            lut = (torch.cumsum(histo, 0) + (step // 2)) // step
            lut = torch.cat([torch.zeros(1), lut[:-1]]).clamp(0, 255)
            out[i, j] = torch.gather(lut, 0, im.flatten().long()).reshape(im.shape) / 255
    return out


# ------------------------------------------------------------------ CLAHE
def _tiles_and_pad(x: torch.Tensor, grid_size):
    b, c, h, w = x.shape
    gh, gw = grid_size
    th, tw = math.ceil(h / gh), math.ceil(w / gw)
    th += th % 2
    tw += tw % 2
    pv, ph = th * gh - h, tw * gw - w
    if pv > h or ph > w:
        raise ValueError("Cannot compute tiles on the image according to the given grid size")
    xp = F.pad(x, [0, ph, 0, pv], mode="reflect") if (pv > 0 or ph > 0) else x
    return xp, th, tw


def compute_luts(x: torch.Tensor, clip_limit: float, grid_size) -> torch.Tensor:
    """-> (B, gh, gw, C, 256) fp32 integer-valued LUTs."""
    b, c, h, w = x.shape
    gh, gw = grid_size
    xp, th, tw = _tiles_and_pad(x, grid_size)
    tiles = xp.unfold(2, th, th).unfold(3, tw, tw).permute(0, 2, 3, 1, 4, 5).contiguous()  # B,gh,gw,C,th,tw
    pixels = th * tw
    flat = tiles.view(-1, pixels)
    histos = torch.stack([torch.histc(t, bins=256, min=0, max=1) for t in flat])
    if clip_limit > 0.0:
        max_val = max(clip_limit * pixels // 256, 1)
        histos.clamp_(max=max_val)
        clipped = pixels - histos.sum(1)
        residual = torch.remainder(clipped, 256)
        redist = (clipped - residual).div(256)
        histos += redist[:, None]
        histos += (torch.arange(256)[None, :] < residual[:, None]).to(histos.dtype)
    lut_scale = 255 / pixels
    luts = (torch.cumsum(histos, 1) * lut_scale).clamp(0, 255).floor()
    return luts.view(b, gh, gw, c, 256)


def equalize_clahe(x: torch.Tensor, clip_limit: float = 40.0, grid_size=(8, 8)) -> torch.Tensor:
    """(B,C,H,W) fp32 in [0,1] -> same shape."""
    b, c, h, w = x.shape
    gh, gw = grid_size
    xp, th, tw = _tiles_and_pad(x, grid_size)
    luts = compute_luts(x, clip_limit, grid_size)  # B,gh,gw,C,256
    hh, hw = th // 2, tw // 2
    it = xp.unfold(2, hh, hh).unfold(3, hw, hw).permute(0, 2, 3, 1, 4, 5).contiguous()  # B,2gh,2gw,C,hh,hw
    g2h, g2w = 2 * gh, 2 * gw

    # neighbouring hist-tile indices of every half-tile along one axis
    def nb(n_half, n_tiles):
        lo = torch.empty(n_half, dtype=torch.long)
        hi = torch.empty(n_half, dtype=torch.long)
        for j in range(n_half):
            if j == 0:
                lo[j] = hi[j] = 0
            elif j == n_half - 1:
                lo[j] = hi[j] = n_tiles - 1
            else:
                lo[j] = (j - 1) // 2
                hi[j] = (j - 1) // 2 + 1
        return lo, hi

    jlo, jhi = nb(g2h, gh)
    ilo, ihi = nb(g2w, gw)
    idx = (it * 255).long().clamp_(0, 255).flatten(-2, -1)  # B,2gh,2gw,C,hh*hw

    def look(jsel, isel):
        m = luts[:, jsel][:, :, isel]  # B,2gh,2gw,C,256
        return torch.gather(m, 4, idx).to(it.dtype).reshape(it.shape)

    tl, tr, bl, br = look(jlo, ilo), look(jlo, ihi), look(jhi, ilo), look(jhi, ihi)

    # weights: a ramp over one full hist tile (two half tiles), denominator T-1
    ih = torch.arange(2 * hh - 1, -1, -1, dtype=it.dtype).div(2.0 * hh - 1)  # 2*hh
    iw = torch.arange(2 * hw - 1, -1, -1, dtype=it.dtype).div(2.0 * hw - 1)
    wy = torch.zeros(g2h, hh, dtype=it.dtype)
    wx = torch.zeros(g2w, hw, dtype=it.dtype)
    for j in range(1, g2h - 1):
        wy[j] = ih[:hh] if (j % 2 == 1) else ih[hh:]
    for i in range(1, g2w - 1):
        wx[i] = iw[:hw] if (i % 2 == 1) else iw[hw:]
    wyb = wy.view(1, g2h, 1, 1, hh, 1).expand_as(it)
    wxb = wx.view(1, 1, g2w, 1, 1, hw).expand_as(it)
    t = torch.addcmul(tr, wxb, tl - tr)
    bb = torch.addcmul(br, wxb, bl - br)
    eq = torch.addcmul(bb, wyb, t - bb).div(255.0)
    out = eq.permute(0, 3, 1, 4, 2, 5).reshape(b, c, g2h * hh, g2w * hw)
    return out[..., :h, :w]
