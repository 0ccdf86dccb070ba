"""scikit-image call signatures over the same sm_100a kernels.

scikit-image >= 0.26.0 is the reference's second image-processing dependency (reference
pyproject.toml:12, pin uv.lock:619-650).  Its filters delegate to scipy.ndimage, whose Gaussian is the
kornia Gaussian with K = 2*int(truncate*sigma + 0.5) + 1 taps
(site-packages/scipy/ndimage/_filters.py:656-669, 745-747) and whose border modes map onto the ABI's:

    scipy / skimage mode   'nearest'    'reflect'     'mirror'    'constant'   'wrap'
    include/mie.h          REPLICATE    SYMMETRIC     REFLECT     CONSTANT     CIRCULAR

    gaussian(image, sigma, mode=, truncate=)            <- skimage.filters.gaussian
    unsharp_mask(image, radius, amount, preserve_range) <- skimage.filters.unsharp_mask
    median(image, ...)                                  <- skimage.filters.median           (filters.py)
    denoise_nl_means(image, ...)                        <- skimage.restoration.denoise_nl_means (filters.py)

Scale: skimage converts integer images to float with img_as_float (unsigned: v / max; signed: [-1, 1]).
Here integer tensors use this package's mapping x01 = (v - lo) / (hi - lo) with the dtype's full range
(uint8 / uint16: identical to img_as_float; int16: the affine image of skimage's scale) and come back in
the SAME integer dtype unless out_dtype=torch.float32 — the natural contract for 16-bit slices.  The
arithmetic is fp32 (skimage: float64 for integer input): agreement with scipy.ndimage is ~2e-7 abs on
[0,1] data (tests/test_skimage_compat.py).  equalize_adapthist / equalize_hist / denoise_bilateral have
different algorithms from their kornia counterparts and are not built.
"""
from __future__ import annotations

import torch

from ._ffi import BORDER, DTYPE_CODE, as_planes, check, lib, require_cuda, stream_ptr, value_range_of
from .filters import MAX_TAPS, _out_like, denoise_nl_means, gaussian_blur2d, get_gaussian_kernel1d, median  # noqa: F401

__all__ = ["gaussian", "unsharp_mask", "median", "denoise_nl_means", "SCIPY_MODES"]

SCIPY_MODES = {"nearest": "replicate", "reflect": "symmetric", "mirror": "reflect", "constant": "constant",
               "wrap": "circular"}


def _taps(sigma: float, truncate: float) -> int:
    if not float(sigma) > 0.0:
        raise ValueError("sigma must be positive")
    k = 2 * int(float(truncate) * float(sigma) + 0.5) + 1
    if k > MAX_TAPS:
        raise ValueError(f"sigma * truncate needs {k} taps; up to {MAX_TAPS} are supported")
    return k


def _mode(mode: str, cval: float) -> str:
    if mode not in SCIPY_MODES:
        raise ValueError(f"mode must be one of {sorted(SCIPY_MODES)}. Got {mode!r}")
    if mode == "constant" and cval != 0:
        raise NotImplementedError("mode='constant' supports cval=0 only")
    return SCIPY_MODES[mode]


def gaussian(image: torch.Tensor, sigma=1, *, mode: str = "nearest", cval=0, preserve_range: bool = False,
             truncate: float = 4.0, channel_axis=None, out=None, value_range=None, out_dtype=None) -> torch.Tensor:
    """skimage.filters.gaussian on 2-D planes ((H,W), (C,H,W), (B,C,H,W): every plane on its own).
    `sigma`: scalar or (sigma_y, sigma_x).  `preserve_range` is accepted for signature compatibility:
    integer tensors are returned in their own dtype and range either way (see the module docstring)."""
    if channel_axis is not None:
        raise NotImplementedError("planes are filtered separately; channel_axis is not supported")
    if out is not None:
        raise NotImplementedError("out= is not supported")
    sy, sx = (sigma if isinstance(sigma, (tuple, list)) else (sigma, sigma))
    return gaussian_blur2d(image, (_taps(sy, truncate), _taps(sx, truncate)), (float(sy), float(sx)), _mode(mode, cval),
                           value_range=value_range, out_dtype=out_dtype)


def unsharp_mask(image: torch.Tensor, radius=1.0, amount=1.0, preserve_range: bool = False, *, channel_axis=None,
                 value_range=None, out_dtype=None) -> torch.Tensor:
    """skimage.filters.unsharp_mask: image + amount * (image - gaussian(image, sigma=radius, mode='reflect')),
    clipped to [0, 1] on the normalised scale unless preserve_range (integer outputs saturate at the dtype's
    range in both cases)."""
    if channel_axis is not None:
        raise NotImplementedError("planes are filtered separately; channel_axis is not supported")
    require_cuda(image)
    k = _taps(radius, 4.0)
    x, n, h, w = as_planes(image)
    lo, hi = value_range_of(x, value_range)
    dst = _out_like(x, out_dtype)
    wk = get_gaussian_kernel1d(k, float(radius))
    with torch.cuda.device(x.device):
        check(lib().mie_unsharp_amount(x.data_ptr(), dst.data_ptr(), DTYPE_CODE[x.dtype], DTYPE_CODE[dst.dtype], n, h, w,
                                       h * w, w, h * w, w, wk.ctypes.data, k, wk.ctypes.data, k, BORDER["symmetric"],
                                       float(amount), 0 if preserve_range else 1, lo, hi, stream_ptr(x.device)))
    return dst
