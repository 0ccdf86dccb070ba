"""numpy front-end of the CPU ORACLE (oracle/mie_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the shipped package never does.

PARITY UNPINNED (see mie_oracle.c header and DESIGN.md §3): the reference holds no
implementation or golden vectors; kornia 0.8.2 / scikit-image 0.26.0 semantics are
restated from their published algorithms and pinned against cv2, scipy.ndimage,
torchvision and oracle/kornia_twin.py where those coincide.

All functions take / return numpy arrays of shape (..., H, W); leading dimensions
are independent planes.
"""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from build import build_oracle  # noqa: E402

_lib = None

BORDERS = {"constant": 0, "reflect": 1, "replicate": 2, "circular": 3, "symmetric": 4}
_DT = {np.dtype(np.uint8): 0, np.dtype(np.uint16): 1, np.dtype(np.int16): 2, np.dtype(np.float32): 3}
_RANGES = {np.dtype(np.uint8): (0.0, 255.0), np.dtype(np.uint16): (0.0, 65535.0), np.dtype(np.int16): (-32768.0, 32767.0)}


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build_oracle())
        for name in dir(_Sig):
            if name.startswith("orc_"):
                fn = getattr(_lib, name)
                fn.argtypes, fn.restype = getattr(_Sig, name)
    return _lib


_p = C.c_void_p
_i64, _i, _f, _d = C.c_int64, C.c_int, C.c_float, C.c_double


class _Sig:
    orc_to01 = ([_p, _i, _i64, _f, _f, _p], None)
    orc_from01 = ([_p, _i, _i64, _f, _f, _p], None)
    orc_gaussian2d = ([_p, _p, _i64, _i, _i, _p, _i, _p, _i, _i, _i], _i)
    orc_gaussian2d_ex = ([_p, _p, _i64, _i, _i, _p, _i, _p, _i, _i, _i, _f, _i], _i)
    orc_clahe_hist_kornia = ([_p, _i64, _i, _i, _i, _i, _p], _i)
    orc_clahe_luts_from_hist_kornia = ([_p, _i64, _i, _i, _d, _p], _i)
    orc_clahe_apply_kornia = ([_p, _p, _i64, _i, _i, _i, _i, _p], _i)
    orc_clahe_hist_opencv_u8 = ([_p, _i64, _i, _i, _i, _i, _p], _i)
    orc_clahe_luts_from_hist_opencv = ([_p, _i64, _i, _i, _d, _p], _i)
    orc_clahe_apply_opencv_u8 = ([_p, _p, _i64, _i, _i, _i, _i, _p], _i)
    orc_clahe_opencv_u16 = ([_p, _p, _i64, _i, _i, _i, _i, _d, _p], _i)
    orc_nlm_fast = ([_p, _p, _i64, _i, _i, _i, _i, _d, _d], _i)
    orc_nlm_slow = ([_p, _p, _i64, _i, _i, _i, _i, _d, _d], _i)
    orc_nlm_patch_weights = ([_i, _d, _p], _i)
    orc_median2d = ([_p, _p, _i64, _i, _i, _i, _i, _i], _i)
    orc_median3d = ([_p, _p, _i, _i, _i, _p, _p, _i], _i)
    orc_exp = ([_p, _p, _i64], None)
    orc_exp2n = ([_p, _p, _i64], None)
    orc_bilateral = ([_p, _p, _i64, _i, _i, _p, _i, _i, _f, _i], _i)
    orc_equalize = ([_p, _p, _i64, _i, _i], _i)
    orc_sk_adapthist_grey = ([_p, _i, _i, _i, _p], _i)
    orc_sk_clahe = ([_p, _i, _i, _i, _i, _d, _i, _p, _p], _i)
    orc_sk_rescale01 = ([_p, _i, _i, _i, _p], _i)
    orc_sk_equalize_hist = ([_p, _i, _i, _i, _p], _i)
    orc_sk_bilateral = ([_p, _i, _i, _i, _i, _i, _i, _d, _p, _p, _p], _i)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _planes(a: np.ndarray, dtype=None):
    a = np.ascontiguousarray(a, dtype=dtype)
    if a.ndim < 2:
        raise ValueError("expected (..., H, W)")
    h, w = a.shape[-2:]
    n = int(np.prod(a.shape[:-2], dtype=np.int64)) if a.ndim > 2 else 1
    return a, n, h, w


def _check(rc: int):
    if rc == -5:
        raise ValueError("grid_size entries must be positive")
    if rc == -6:
        raise ValueError("Cannot compute tiles on the image according to the given grid size")
    if rc != 0:
        raise RuntimeError(f"oracle error {rc}")


# ------------------------------------------------------------------ pixel mapping
def default_range(dtype):
    return _RANGES[np.dtype(dtype)]


def to01(a: np.ndarray, value_range=None) -> np.ndarray:
    a = np.ascontiguousarray(a)
    if a.dtype == np.float32:
        return a.copy()
    lo, hi = value_range if value_range is not None else default_range(a.dtype)
    out = np.empty(a.shape, np.float32)
    lib().orc_to01(_ptr(a), _DT[a.dtype], a.size, lo, hi, _ptr(out))
    return out


def from01(y: np.ndarray, dtype, value_range=None) -> np.ndarray:
    y = np.ascontiguousarray(y, np.float32)
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return y.copy()
    lo, hi = value_range if value_range is not None else default_range(dtype)
    out = np.empty(y.shape, dtype)
    lib().orc_from01(_ptr(y), _DT[dtype], y.size, lo, hi, _ptr(out))
    return out


# ------------------------------------------------------------------ Gaussian / unsharp
def gaussian_kernel1d(kernel_size: int, sigma: float) -> np.ndarray:
    """kornia.filters.get_gaussian_kernel1d: x = arange(K) - K//2 (+0.5 if even),
    exp(-x^2/(2 sigma^2)) normalised.  Evaluated in float64, rounded once to fp32."""
    k = int(kernel_size)
    x = np.arange(k, dtype=np.float64) - k // 2
    if k % 2 == 0:
        x = x + 0.5
    g = np.exp(-(x * x) / (2.0 * float(sigma) ** 2))
    return (g / g.sum()).astype(np.float32)


def _pair(v):
    if isinstance(v, (tuple, list)):
        return v[0], v[1]
    return v, v


def _sep(x01, kernel_size, sigma, border_type, unsharp):
    x, n, h, w = _planes(x01, np.float32)
    ky, kx = _pair(kernel_size)
    sy, sx = _pair(sigma)
    wx, wy = gaussian_kernel1d(kx, sx), gaussian_kernel1d(ky, sy)
    out = np.empty_like(x)
    rc = lib().orc_gaussian2d(_ptr(x), _ptr(out), n, h, w, _ptr(wx), kx, _ptr(wy), ky, BORDERS[border_type], unsharp)
    _check(rc)
    return out


def skimage_gaussian(x01, sigma=1.0, mode="nearest", truncate=4.0):
    """skimage.filters.gaussian -> scipy.ndimage.gaussian_filter on (..., H, W) float planes:
    radius int(truncate*sigma + 0.5) (site-packages/scipy/ndimage/_filters.py:745-747), weights
    exp(-0.5 x^2/sigma^2) normalised (:656-669) — the kornia formula with K = 2 radius + 1."""
    k = 2 * int(truncate * float(sigma) + 0.5) + 1
    return _sep(x01, k, sigma, SCIPY_MODES[mode], 0)


def skimage_unsharp_mask(x01, radius=1.0, amount=1.0, preserve_range=False):
    """skimage.filters.unsharp_mask on [0,1] planes: x + amount (x - gaussian(x, sigma=radius,
    mode='reflect')), clipped to [0,1] unless preserve_range."""
    x, n, h, w = _planes(x01, np.float32)
    k = 2 * int(4.0 * float(radius) + 0.5) + 1
    wk = gaussian_kernel1d(k, radius)
    out = np.empty_like(x)
    _check(lib().orc_gaussian2d_ex(_ptr(x), _ptr(out), n, h, w, _ptr(wk), k, _ptr(wk), k, BORDERS["symmetric"], 1,
                                   float(amount), 0 if preserve_range else 1))
    return out


SCIPY_MODES = {"nearest": "replicate", "reflect": "symmetric", "mirror": "reflect", "constant": "constant",
               "wrap": "circular"}


def gaussian_blur2d(x01, kernel_size, sigma, border_type="reflect"):
    return _sep(x01, kernel_size, sigma, border_type, 0)


def unsharp_mask(x01, kernel_size, sigma, border_type="reflect"):
    return _sep(x01, kernel_size, sigma, border_type, 1)


# ------------------------------------------------------------------ CLAHE (kornia semantics)
def kornia_tile_size(h, w, grid_size):
    gh, gw = grid_size
    th, tw = -(-h // gh), -(-w // gw)
    return th + (th & 1), tw + (tw & 1)


def clahe_hist(x01, grid_size=(8, 8)) -> np.ndarray:
    x, n, h, w = _planes(x01, np.float32)
    gh, gw = grid_size
    if gh <= 0 or gw <= 0:
        raise ValueError("grid_size entries must be positive")
    hist = np.empty((n, gh, gw, 256), np.uint32)
    _check(lib().orc_clahe_hist_kornia(_ptr(x), n, h, w, gh, gw, _ptr(hist)))
    return hist.reshape(x.shape[:-2] + (gh, gw, 256))


def clahe_luts_from_hist(hist, tile_size, clip_limit) -> np.ndarray:
    hist = np.ascontiguousarray(hist, np.uint32)
    luts = np.empty(hist.shape, np.uint8)
    th, tw = tile_size
    _check(lib().orc_clahe_luts_from_hist_kornia(_ptr(hist), hist.size // 256, th, tw, float(clip_limit), _ptr(luts)))
    return luts


def clahe_luts(x01, clip_limit=40.0, grid_size=(8, 8)) -> np.ndarray:
    h, w = x01.shape[-2:]
    return clahe_luts_from_hist(clahe_hist(x01, grid_size), kornia_tile_size(h, w, grid_size), clip_limit)


def clahe_apply(x01, luts, grid_size=(8, 8)) -> np.ndarray:
    x, n, h, w = _planes(x01, np.float32)
    gh, gw = grid_size
    luts = np.ascontiguousarray(luts, np.uint8)
    assert luts.size == n * gh * gw * 256
    out = np.empty_like(x)
    _check(lib().orc_clahe_apply_kornia(_ptr(x), _ptr(out), n, h, w, gh, gw, _ptr(luts)))
    return out


def equalize_clahe(x01, clip_limit=40.0, grid_size=(8, 8)) -> np.ndarray:
    return clahe_apply(x01, clahe_luts(x01, clip_limit, grid_size), grid_size)


# ------------------------------------------------------------------ CLAHE (OpenCV semantics, uint8)
def opencv_tile_size(h, w, grid_size):
    """cv::CLAHE pads both axes by tiles - dim % tiles when either is not divisible (a divisible axis then grows by
    a full `tiles` pixels)."""
    gh, gw = grid_size
    if h % gh == 0 and w % gw == 0:
        return h // gh, w // gw
    return (h + gh - h % gh) // gh, (w + gw - w % gw) // gw


def opencv_clahe_hist(img, grid_size=(8, 8)) -> np.ndarray:
    x, n, h, w = _planes(img, np.uint8)
    gh, gw = grid_size
    hist = np.empty((n, gh, gw, 256), np.uint32)
    _check(lib().orc_clahe_hist_opencv_u8(_ptr(x), n, h, w, gh, gw, _ptr(hist)))
    return hist.reshape(x.shape[:-2] + (gh, gw, 256))


def opencv_clahe_luts(img, clip_limit=40.0, grid_size=(8, 8)) -> np.ndarray:
    h, w = img.shape[-2:]
    hist = opencv_clahe_hist(img, grid_size)
    luts = np.empty(hist.shape, np.uint8)
    th, tw = opencv_tile_size(h, w, grid_size)
    _check(lib().orc_clahe_luts_from_hist_opencv(_ptr(hist), hist.size // 256, th, tw, float(clip_limit), _ptr(luts)))
    return luts


def opencv_clahe16(img, clip_limit=40.0, grid_size=(8, 8), return_luts=False):
    """cv2.createCLAHE(clipLimit, tileGridSize=(gw, gh)).apply(img) for uint16 (65 536 bins)."""
    x, n, h, w = _planes(img, np.uint16)
    gh, gw = grid_size
    out = np.empty_like(x)
    luts = np.empty((n, gh, gw, 65536), np.uint16) if return_luts else None
    _check(lib().orc_clahe_opencv_u16(_ptr(x), _ptr(out), n, h, w, gh, gw, float(clip_limit),
                                      _ptr(luts) if return_luts else None))
    out = out.reshape(np.asarray(img).shape)
    return (out, luts) if return_luts else out


def opencv_clahe(img, clip_limit=40.0, grid_size=(8, 8)) -> np.ndarray:
    """cv2.createCLAHE(clipLimit, tileGridSize=(gw, gh)).apply(img) for uint8; grid_size is (rows, cols)."""
    if np.asarray(img).dtype == np.uint16:
        return opencv_clahe16(img, clip_limit, grid_size)
    x, n, h, w = _planes(img, np.uint8)
    gh, gw = grid_size
    luts = opencv_clahe_luts(x, clip_limit, grid_size)
    out = np.empty_like(x)
    _check(lib().orc_clahe_apply_opencv_u8(_ptr(x), _ptr(out), n, h, w, gh, gw, _ptr(luts)))
    return out


# ------------------------------------------------------------------ non-local means (skimage fast mode)
def denoise_nl_means(x01, patch_size=7, patch_distance=11, h=0.1, sigma=0.0) -> np.ndarray:
    """skimage.restoration.denoise_nl_means(fast_mode=True) on (..., H, W) planes of [0,1] data; float64."""
    x, n, hh, ww = _planes(np.asarray(x01), np.float64)
    out = np.empty_like(x)
    rc = lib().orc_nlm_fast(_ptr(x), _ptr(out), n, hh, ww, int(patch_size), int(patch_distance), float(h), float(sigma))
    if rc:
        raise ValueError("invalid non-local-means parameters (patch / distance too large for the image?)")
    return out


def nlm_fast_literal(image, patch_size=7, patch_distance=11, h=0.1, sigma=0.0) -> np.ndarray:
    """Literal numpy transcription of skimage's _fast_nl_means_denoising_2d loops [RECALLED] (integral
    image per shift, symmetric accumulation with alpha = 0.5 on the t_col == 0 column), single 2-D image,
    float64.  Slow: for pinning orc_nlm_fast on small images only."""
    image = np.asarray(image, np.float64)
    s = patch_size + (1 if patch_size % 2 == 0 else 0)
    offset, d = s // 2, patch_distance
    pad = offset + d + 1
    padded = np.pad(image, pad, mode="reflect")
    n_row, n_col = padded.shape
    result = np.zeros_like(padded)
    weights = np.zeros_like(padded)
    var = 2.0 * sigma * sigma
    h2s2 = h * h * s * s
    for t_row in range(-d, d + 1):
        row_start, row_end = max(offset, offset - t_row), min(n_row - offset, n_row - offset - t_row)
        for t_col in range(0, d + 1):
            alpha = 0.5 if (t_col == 0 and t_row != 0) else 1.0
            col_start, col_end = max(offset, offset - t_col), min(n_col - offset, n_col - offset - t_col)
            integral = np.zeros_like(padded)
            r0, r1 = max(1, -t_row), min(n_row, n_row - t_row)
            c0, c1 = max(1, -t_col), min(n_col, n_col - t_col)
            diff = (padded[r0:r1, c0:c1] - padded[r0 + t_row:r1 + t_row, c0 + t_col:c1 + t_col]) ** 2 - var
            integral[r0:r1, c0:c1] = np.cumsum(np.cumsum(diff, axis=0), axis=1)
            for row in range(row_start, row_end):
                for col in range(col_start, col_end):
                    dist = (integral[row + offset, col + offset] + integral[row - offset, col - offset]
                            - integral[row - offset, col + offset] - integral[row + offset, col - offset])
                    dist = max(dist, 0.0) / h2s2
                    if dist > 5.0:
                        continue
                    wgt = alpha * np.exp(-dist)
                    weights[row, col] += wgt
                    weights[row + t_row, col + t_col] += wgt
                    result[row, col] += wgt * padded[row + t_row, col + t_col]
                    result[row + t_row, col + t_col] += wgt * padded[row, col]
    out = result[pad:-pad, pad:-pad] / weights[pad:-pad, pad:-pad]
    return out


def denoise_nl_means_slow(x01, patch_size=7, patch_distance=11, h=0.1, sigma=0.0) -> np.ndarray:
    """skimage.restoration.denoise_nl_means(fast_mode=False) on (..., H, W) planes of [0,1] data; float64."""
    x, n, hh, ww = _planes(np.asarray(x01), np.float64)
    out = np.empty_like(x)
    rc = lib().orc_nlm_slow(_ptr(x), _ptr(out), n, hh, ww, int(patch_size), int(patch_distance), float(h), float(sigma))
    if rc:
        raise ValueError("invalid non-local-means parameters (patch too large for the image?)")
    return out


def nlm_slow_literal(image, patch_size=7, patch_distance=11, h=0.1, sigma=0.0) -> np.ndarray:
    """Literal transcription of skimage's _nl_means_denoising_2d / patch_distance_2d loops [RECALLED] (np.pad by the
    patch radius, meshgrid Gaussian patch weights, search window clipped at the image, cut-off test before every
    patch row), single 2-D image, float64.  Slow: for pinning orc_nlm_slow on small images only."""
    image = np.asarray(image, np.float64)
    s = patch_size + (1 if patch_size % 2 == 0 else 0)
    d = patch_distance
    n_row, n_col = image.shape
    offset = s // 2
    padded = np.ascontiguousarray(np.pad(image, ((offset, offset), (offset, offset)), mode="reflect"))
    result = np.empty_like(image)
    A = (s - 1.0) / 4.0
    range_vals = np.arange(-offset, offset + 1, dtype=np.float64)
    xg_row, xg_col = np.meshgrid(range_vals, range_vals, indexing="ij")
    w = np.ascontiguousarray(np.exp(-(xg_row * xg_row + xg_col * xg_col) / (2 * A * A)))
    w *= 1.0 / (np.sum(w) * h * h)
    var = 2.0 * sigma * sigma

    def patch_distance_2d(p1, p2):
        distance = 0.0
        for i in range(s):
            if distance > 5.0:
                return 0.0
            for j in range(s):
                tmp_diff = p1[i, j] - p2[i, j]
                distance += w[i, j] * (tmp_diff * tmp_diff - var)
        return np.exp(-max(0.0, distance))

    for row in range(n_row):
        i_start = row - min(d, row)
        i_end = row + min(d + 1, n_row - row)
        for col in range(n_col):
            new_value = 0.0
            weight_sum = 0.0
            j_start = col - min(d, col)
            j_end = col + min(d + 1, n_col - col)
            central_patch = padded[row:row + s, col:col + s]
            for i in range(i_start, i_end):
                for j in range(j_start, j_end):
                    weight = patch_distance_2d(central_patch, padded[i:i + s, j:j + s])
                    weight_sum += weight
                    new_value += weight * padded[i + offset, j + offset]
            result[row, col] = new_value / weight_sum
    return result


# ------------------------------------------------------------------ median / bilateral / equalize
def median_blur(a, kernel_size, border_type="constant") -> np.ndarray:
    """kornia.filters.median_blur (zero padding) on (..., H, W); any dtype, same dtype out."""
    a = np.ascontiguousarray(a)
    x, n, h, w = _planes(a, np.float64)
    ky, kx = _pair(kernel_size)
    out = np.empty_like(x)
    _check(lib().orc_median2d(_ptr(x), _ptr(out), n, h, w, ky, kx, BORDERS[border_type]))
    return out.astype(a.dtype)


def median3d(vol, mode="nearest", halo_lo=None, halo_hi=None) -> np.ndarray:
    """skimage.filters.median / scipy.ndimage.median_filter 3x3x3 on a (D, H, W) volume."""
    vol = np.ascontiguousarray(vol)
    d, h, w = vol.shape
    x = vol.astype(np.float64)
    lo = None if halo_lo is None else np.ascontiguousarray(halo_lo, np.float64)
    hi = None if halo_hi is None else np.ascontiguousarray(halo_hi, np.float64)
    out = np.empty_like(x)
    border = {"nearest": 2, "constant": 0}[mode]
    _check(lib().orc_median3d(_ptr(x), _ptr(out), d, h, w, None if lo is None else _ptr(lo),
                              None if hi is None else _ptr(hi), border))
    return out.astype(vol.dtype)


def mie_exp2n(t) -> np.ndarray:
    """2^t for t <= 0 as the bilateral filter evaluates it (oracle/mie_oracle.c:mie_exp2n)."""
    t = np.ascontiguousarray(t, np.float32)
    out = np.empty_like(t)
    lib().orc_exp2n(_ptr(t), _ptr(out), t.size)
    return out


def mie_exp(a) -> np.ndarray:
    a = np.ascontiguousarray(a, np.float32)
    out = np.empty_like(a)
    lib().orc_exp(_ptr(a), _ptr(out), a.size)
    return out


def bilateral_blur(x01, kernel_size, sigma_color, sigma_space, border_type="reflect") -> np.ndarray:
    x, n, h, w = _planes(x01, np.float32)
    ky, kx = _pair(kernel_size)
    sy, sx = _pair(sigma_space)
    wsp = np.ascontiguousarray((gaussian_kernel1d(ky, sy)[:, None] * gaussian_kernel1d(kx, sx)[None, :]).astype(np.float32))
    out = np.empty_like(x)
    _check(lib().orc_bilateral(_ptr(x), _ptr(out), n, h, w, _ptr(wsp), ky, kx, float(sigma_color), BORDERS[border_type]))
    return out


def equalize(x01) -> np.ndarray:
    x, n, h, w = _planes(x01, np.float32)
    out = np.empty_like(x)
    _check(lib().orc_equalize(_ptr(x), _ptr(out), n, h, w))
    return out


# ------------------------------------------------------------------ chain (BASELINE.json config 2)
def chain_gauss_clahe_unsharp(img, denoise_kernel=9, denoise_sigma=1.0, clip_limit=2.0, grid_size=(8, 8),
                              sharpen_kernel=9, sharpen_sigma=1.0, border_type="reflect", value_range=None,
                              out_dtype=None, return_stages=False):
    """Gaussian denoise -> CLAHE -> unsharp mask on [0,1] fp32, quantised once at the end."""
    img = np.ascontiguousarray(img)
    out_dtype = np.dtype(out_dtype) if out_dtype is not None else img.dtype
    x01 = to01(img, value_range)
    g = gaussian_blur2d(x01, denoise_kernel, denoise_sigma, border_type)
    luts = clahe_luts(g, clip_limit, grid_size)
    c = clahe_apply(g, luts, grid_size)
    u = unsharp_mask(c, sharpen_kernel, sharpen_sigma, border_type)
    out = from01(u, out_dtype, value_range)
    if return_stages:
        return out, {"x01": x01, "gauss": g, "luts": luts, "clahe": c, "unsharp": u}
    return out


# ---------------------------------------------------------------------------- sewar-style metrics (F4)
# numpy / scipy float64 restatement of sewar 0.4.6 full_ref.{mse,rmse,psnr,ssim} (reference
# pyproject.toml:13, uv.lock:692-700; package not installable here -> RECALLED, parity unpinned):
#   mse  = mean((GT.astype(float64) - P.astype(float64))**2)
#   psnr = 10 log10(MAX**2 / mse), inf if mse == 0, MAX = np.iinfo(GT.dtype).max by default
#   ssim : per channel, win = uniform ws x ws / ws**2 (sewar's default fltr_specs), filter2(..., 'valid')
#          = scipy.signal.convolve2d with the rotated window; mu, sigma from E[x^2] - mu^2;
#          returns (mean ssim_map, mean cs_map), averaged over channels.
# Inputs here are (..., H, W): every leading plane is a "channel".
def _sewar_planes(a):
    a = np.asarray(a)
    return a.reshape((-1,) + a.shape[-2:])


def sewar_mse(GT, P) -> float:
    return float(np.mean((np.asarray(GT).astype(np.float64) - np.asarray(P).astype(np.float64)) ** 2))


def sewar_rmse(GT, P) -> float:
    return float(np.sqrt(sewar_mse(GT, P)))


def sewar_psnr(GT, P, MAX=None) -> float:
    if MAX is None:
        MAX = np.iinfo(np.asarray(GT).dtype).max
    m = sewar_mse(GT, P)
    return float("inf") if m == 0.0 else float(10 * np.log10(MAX ** 2 / m))


def sewar_ssim(GT, P, ws=11, K1=0.01, K2=0.03, MAX=None):
    from scipy.signal import convolve2d

    if MAX is None:
        MAX = np.iinfo(np.asarray(GT).dtype).max
    C1, C2 = (K1 * MAX) ** 2, (K2 * MAX) ** 2
    win = np.ones((ws, ws)) / ws ** 2
    ss, cs = [], []
    for g, p in zip(_sewar_planes(GT).astype(np.float64), _sewar_planes(P).astype(np.float64)):
        f = lambda x: convolve2d(x, np.rot90(win, 2), mode="valid")  # noqa: E731  (sewar.utils.filter2)
        mu1, mu2 = f(g), f(p)
        gt_sum_sq, p_sum_sq, gt_p_sum_mul = mu1 * mu1, mu2 * mu2, mu1 * mu2
        s_gt, s_p, s_gp = f(g * g) - gt_sum_sq, f(p * p) - p_sum_sq, f(g * p) - gt_p_sum_mul
        ssim_map = ((2 * gt_p_sum_mul + C1) * (2 * s_gp + C2)) / ((gt_sum_sq + p_sum_sq + C1) * (s_gt + s_p + C2))
        cs_map = (2 * s_gp + C2) / (s_gt + s_p + C2)
        ss.append(np.mean(ssim_map))
        cs.append(np.mean(cs_map))
    return float(np.mean(ss)), float(np.mean(cs))


# ------------------------------------------------------------------ scikit-image exposure / restoration (RECALLED)
SK_MODES = {"constant": 0, "edge": 1, "symmetric": 2, "reflect": 3, "wrap": 4}


def sk_adapthist_kernel(shape, kernel_size):
    if kernel_size is None:
        return [max(s // 8, 1) for s in shape]
    if isinstance(kernel_size, (int, float)):
        return [int(kernel_size)] * len(shape)
    if len(kernel_size) != len(shape):
        raise ValueError(f"Incorrect value of `kernel_size`: {kernel_size}")
    return [int(k) for k in kernel_size]


def sk_equalize_adapthist(image, kernel_size=None, clip_limit=0.01, nbins=256, return_stages=False):
    """skimage.exposure.equalize_adapthist on ONE 2-D image (per-pixel C restatement; see mie_oracle.c)."""
    x = np.ascontiguousarray(image)
    if x.ndim != 2 or x.dtype not in _DT:
        raise ValueError("expected one 2-D image of dtype uint8 / uint16 / int16 / float32")
    h, w = x.shape
    kr, kc = sk_adapthist_kernel(x.shape, kernel_size)
    grey = np.empty((h, w), np.uint16)
    _check(lib().orc_sk_adapthist_grey(_ptr(x), _DT[x.dtype], h, w, _ptr(grey)))
    c = np.empty((h, w), np.uint16)
    nbr, nbc = -(-h // kr), -(-w // kc)
    maps = np.empty((nbr, nbc, nbins), np.int64)
    _check(lib().orc_sk_clahe(_ptr(grey), h, w, kr, kc, float(clip_limit), int(nbins), _ptr(c), _ptr(maps)))
    f32 = x.dtype == np.float32
    out = np.empty((h, w), np.float32 if f32 else np.float64)
    _check(lib().orc_sk_rescale01(_ptr(c), h, w, int(f32), _ptr(out)))
    if return_stages:
        return out, {"grey": grey, "clahe": c, "maps": maps}
    return out


def sk_equalize_hist(image):
    x = np.ascontiguousarray(image)
    if x.ndim != 2 or x.dtype not in _DT or x.dtype == np.float32:
        raise ValueError("expected one 2-D image of dtype uint8 / uint16 / int16")
    out = np.empty(x.shape, np.float64)
    _check(lib().orc_sk_equalize_hist(_ptr(x), _DT[x.dtype], x.shape[0], x.shape[1], _ptr(out)))
    return out


def sk_bilateral_luts(bins, sigma_color, max_value, win_size, sigma_spatial):
    """The two LUTs upstream builds in Python (numpy exp, float64): colour weights over [0, max_value) and the spatial
    Gaussian of the win_size x win_size window."""
    values = np.linspace(0, max_value, bins, endpoint=False)
    color = np.exp(-0.5 * (values ** 2 / sigma_color ** 2))
    ext = (win_size - 1) // 2
    g = np.arange(-ext, ext + 1)
    rr, cc = np.meshgrid(g, g, indexing="ij")
    spatial = np.exp(-0.5 * (np.hypot(rr, cc) ** 2 / sigma_spatial ** 2)).ravel()
    return np.ascontiguousarray(color), np.ascontiguousarray(spatial)


def sk_denoise_bilateral(image, win_size=None, sigma_color=None, sigma_spatial=1, bins=10000, mode="constant", cval=0):
    """skimage.restoration.denoise_bilateral on ONE 2-D integer image (per-pixel C restatement)."""
    import math

    x = np.ascontiguousarray(image)
    if x.ndim != 2 or x.dtype not in _DT or x.dtype == np.float32:
        raise ValueError("expected one 2-D image of dtype uint8 / uint16 / int16")
    info = np.iinfo(x.dtype)
    if x.dtype.kind == "u":
        f = x.astype(np.float64) / float(info.max)
    else:
        f = (x.astype(np.float64) * 2.0 + 1.0) / (float(info.max) - float(info.min))
    sigma_color = sigma_color or f.std()
    if win_size is None:
        win_size = max(5, 2 * int(math.ceil(3 * sigma_spatial)) + 1)
    mn, mx = f.min(), f.max()
    out = np.empty(x.shape, np.float64)
    if mn == mx:
        out[...] = mn
        return out
    max_value = mx - mn if mn < 0 else mx
    color, spatial = sk_bilateral_luts(bins, sigma_color, max_value, win_size, sigma_spatial)
    _check(lib().orc_sk_bilateral(_ptr(x), _DT[x.dtype], x.shape[0], x.shape[1], int(win_size), int(bins), SK_MODES[mode],
                                  float(cval), _ptr(color), _ptr(spatial), _ptr(out)))
    return out
