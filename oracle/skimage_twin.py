"""numpy "scikit-image twin" — TEST INFRASTRUCTURE ONLY (never imported by the package).

scikit-image 0.26.0 (reference pyproject.toml:12, pin uv.lock:619-650) cannot be installed in this image
(no network; the lock targets cp314).  This module restates, array-level and in numpy as upstream does, the three
functions of SURVEY.md §8(f) F3 that differ algorithmically from their kornia counterparts:

    skimage.exposure.equalize_hist        (np.interp of the image's own CDF)
    skimage.exposure.equalize_adapthist   (CLAHE on 2^14 grey levels, iterative clip redistribution, block-corner
                                           centred contextual regions, multilinear interpolation, min-max rescale)
    skimage.restoration.denoise_bilateral (colour LUT of `bins` entries, spatial Gaussian LUT, mode='constant')

**RECALLED** (SURVEY.md Appendix B3/B4): written from the published algorithm as the author remembers the 0.2x
sources; the package itself is not on disk, so nothing here is pinned against upstream — `parity unpinned`.
tests/test_live_pins.py compares this twin AND the CUDA path with the real package the day it is importable.
Decisions where the recollection is uncertain are marked `ASSUMED`.

The twin is structurally independent of the per-pixel restatement in mie_oracle.c (orc_sk_*), which is the
formulation the CUDA kernels follow: tests/test_skimage_exposure.py requires the two to agree bit for bit
(float64 / float32 arithmetic in the same order).
"""
from __future__ import annotations

import math

import numpy as np

NR_OF_GRAY = 2 ** 14  # number of grayscale levels to use in CLAHE algorithm


# ------------------------------------------------------------------ dtype helpers (skimage.util.dtype)
def img_as_float(image: np.ndarray) -> np.ndarray:
    """skimage.util.img_as_float for the dtypes of this path: unsigned -> v / max; signed -> (2 v + 1) / (max - min)
    i.e. [-1, 1]; floats unchanged.  Integer input -> float64."""
    image = np.asarray(image)
    if image.dtype.kind == "f":
        return image
    info = np.iinfo(image.dtype)
    out = image.astype(np.float64)
    if image.dtype.kind == "u":
        out /= float(info.max)           # multiply-free: upstream divides (np.divide(image, imax_in))
        return out
    out *= 2.0
    out += 1.0
    out /= float(info.max) - float(info.min)
    return out


def _supported_float_type(dtype) -> np.dtype:
    dtype = np.dtype(dtype)
    if dtype.kind == "f" and dtype.itemsize <= 4:
        return np.dtype(np.float32)
    return np.dtype(np.float64)


def rescale_intensity(image: np.ndarray, out_range) -> np.ndarray:
    """skimage.exposure.rescale_intensity(image, in_range='image', out_range=(omin, omax)) on a float image."""
    imin, imax = image.min(), image.max()
    omin, omax = out_range
    image = np.clip(image, imin, imax)
    if imin != imax:
        image = (image - imin) / (imax - imin)
        return image * (omax - omin) + omin
    return np.clip(image, omin, omax)


# ------------------------------------------------------------------ equalize_hist
def _histogram_int(image: np.ndarray):
    """skimage.exposure.histogram for integer images: one bin per integer value between min and max (`nbins` is
    ignored), bin centres = the values."""
    imin, imax = int(image.min()), int(image.max())
    hist = np.bincount((image.ravel().astype(np.int64) - imin), minlength=imax - imin + 1)
    centers = np.arange(imin, imax + 1)
    return hist, centers


def cumulative_distribution(image: np.ndarray, nbins: int = 256):
    if image.dtype.kind in "ui":
        hist, centers = _histogram_int(image)
    else:
        hist, edges = np.histogram(image.ravel(), bins=nbins, range=(float(image.min()), float(image.max())))
        centers = (edges[:-1] + edges[1:]) / 2.0
    cdf = hist.cumsum()
    cdf = cdf / float(cdf[-1])
    if image.dtype.kind == "f":   # upstream keeps float32 images in float32
        cdf = cdf.astype(_supported_float_type(image.dtype), copy=False)
    return cdf, centers


def equalize_hist(image: np.ndarray, nbins: int = 256) -> np.ndarray:
    """skimage.exposure.equalize_hist(image, nbins, mask=None) on ONE array (any shape: the histogram is global)."""
    image = np.asarray(image)
    cdf, centers = cumulative_distribution(image, nbins)
    out = np.interp(image.ravel(), centers, cdf).reshape(image.shape)
    return out.astype(_supported_float_type(image.dtype), copy=False)


# ------------------------------------------------------------------ equalize_adapthist
def clip_histogram(hist: np.ndarray, clip_limit: int) -> np.ndarray:
    """skimage.exposure._adapthist.clip_histogram: clip, spread the average increment, then hand out the remainder one
    count at a time with a stride that depends on how many bins are still under the limit."""
    hist = hist.copy()
    excess_mask = hist > clip_limit
    excess = hist[excess_mask]
    n_excess = excess.sum() - excess.size * clip_limit
    hist[excess_mask] = clip_limit

    bin_incr = n_excess // hist.size
    upper = clip_limit - bin_incr

    low_mask = hist < upper
    n_excess -= hist[low_mask].size * bin_incr
    hist[low_mask] += bin_incr

    mid_mask = np.logical_and(hist >= upper, hist < clip_limit)
    mid = hist[mid_mask]
    n_excess += mid.sum() - mid.size * clip_limit
    hist[mid_mask] = clip_limit

    while n_excess > 0:
        prev_n_excess = n_excess
        for index in range(hist.size):
            under_mask = hist < clip_limit
            step_size = max(1, np.count_nonzero(under_mask) // n_excess)
            under_mask = under_mask[index::step_size]
            hist[index::step_size][under_mask] += 1
            n_excess -= np.count_nonzero(under_mask)
            if n_excess <= 0:
                break
        if prev_n_excess == n_excess:
            break
    return hist


def map_histogram(hist: np.ndarray, min_val: int, max_val: int, n_pixels: int) -> np.ndarray:
    out = np.cumsum(hist, axis=-1).astype(float)
    out *= (max_val - min_val) / n_pixels
    out += min_val
    np.clip(out, a_min=None, a_max=max_val, out=out)
    return out.astype(int)


def _clahe(image: np.ndarray, kernel_size, clip_limit: float, nbins: int) -> np.ndarray:
    ndim = image.ndim
    dtype = image.dtype
    pad_start = [k // 2 for k in kernel_size]
    pad_end = [(k - s % k) % k + int(np.ceil(k / 2.0)) for k, s in zip(kernel_size, image.shape)]
    image = np.pad(image, [[a, b] for a, b in zip(pad_start, pad_end)], mode="reflect")

    bin_size = 1 + NR_OF_GRAY // nbins
    lut = np.arange(NR_OF_GRAY, dtype=np.min_scalar_type(NR_OF_GRAY))
    lut //= bin_size
    image = lut[image]

    ns_hist = [int(s / k) - 1 for s, k in zip(image.shape, kernel_size)]
    hist_blocks_shape = np.array([ns_hist, kernel_size]).T.flatten()
    hist_blocks_axis_order = np.array([np.arange(0, ndim * 2, 2), np.arange(1, ndim * 2, 2)]).flatten()
    hist_slices = [slice(k // 2, k // 2 + n * k) for k, n in zip(kernel_size, ns_hist)]
    hist_blocks = image[tuple(hist_slices)].reshape(hist_blocks_shape)
    hist_blocks = np.transpose(hist_blocks, axes=hist_blocks_axis_order)
    hist_block_assembled_shape = hist_blocks.shape
    hist_blocks = hist_blocks.reshape((math.prod(ns_hist), -1))

    kernel_elements = math.prod(kernel_size)
    if clip_limit > 0.0:
        clim = int(np.clip(clip_limit * kernel_elements, 1, None))
    else:
        clim = np.iinfo(hist_blocks.dtype).max   # largest possible value: do not clip (plain AHE)

    hist = np.apply_along_axis(np.bincount, -1, hist_blocks, minlength=nbins)
    hist = np.apply_along_axis(clip_histogram, -1, hist, clip_limit=clim)
    hist = map_histogram(hist, 0, NR_OF_GRAY - 1, kernel_elements)
    hist = hist.reshape(hist_block_assembled_shape[:ndim] + (-1,))

    map_array = np.pad(hist, [[1, 1] for _ in range(ndim)] + [[0, 0]], mode="edge")

    ns_proc = [int(s / k) for s, k in zip(image.shape, kernel_size)]
    blocks_shape = np.array([ns_proc, kernel_size]).T.flatten()
    blocks_axis_order = np.array([np.arange(0, ndim * 2, 2), np.arange(1, ndim * 2, 2)]).flatten()
    blocks = image.reshape(blocks_shape)
    blocks = np.transpose(blocks, axes=blocks_axis_order)
    blocks_flattened_shape = blocks.shape
    blocks = np.reshape(blocks, (math.prod(ns_proc), math.prod(blocks.shape[ndim:])))

    coeffs = np.meshgrid(*tuple([np.arange(k) / k for k in kernel_size[::-1]]), indexing="ij")
    coeffs = [np.transpose(c).flatten() for c in coeffs]
    inv_coeffs = [1 - c for c in coeffs]

    result = np.zeros(blocks.shape, dtype=np.float32)
    for edge in np.ndindex(*([2] * ndim)):
        edge_maps = map_array[tuple([slice(e, e + n) for e, n in zip(edge, ns_proc)])]
        edge_maps = edge_maps.reshape((math.prod(ns_proc), -1))
        edge_mapped = np.take_along_axis(edge_maps, blocks, axis=-1)
        edge_coeffs = np.prod([[inv_coeffs, coeffs][e][d] for d, e in enumerate(edge[::-1])], 0)
        result += (edge_mapped * edge_coeffs).astype(result.dtype)

    result = result.astype(dtype)
    result = result.reshape(blocks_flattened_shape)
    blocks_axis_rebuild_order = np.array([np.arange(0, ndim), np.arange(ndim, ndim * 2)]).T.flatten()
    result = np.transpose(result, axes=blocks_axis_rebuild_order)
    result = result.reshape(image.shape)
    unpad = tuple([slice(a, s - b) for a, b, s in zip(pad_start, pad_end, image.shape)])
    return result[unpad]


def adapthist_kernel_size(shape, kernel_size):
    if kernel_size is None:
        return [max(s // 8, 1) for s in shape]
    if isinstance(kernel_size, (int, float)):
        return [int(kernel_size)] * len(shape)
    if len(kernel_size) != len(shape):
        raise ValueError(f"Incorrect value of `kernel_size`: {kernel_size}")
    return [int(k) for k in kernel_size]


def equalize_adapthist(image: np.ndarray, kernel_size=None, clip_limit: float = 0.01, nbins: int = 256,
                       return_stages: bool = False):
    """skimage.exposure.equalize_adapthist on ONE array (2-D here).  Stages (for stage-wise parity tests):
    grey = the 2^14-level image CLAHE sees, clahe = its output (uint16), out = the min-max rescaled float result."""
    image = np.asarray(image)
    float_dtype = _supported_float_type(image.dtype)
    f = img_as_float(image).astype(float_dtype, copy=False)
    grey = np.round(rescale_intensity(f, out_range=(0, NR_OF_GRAY - 1))).astype(np.min_scalar_type(NR_OF_GRAY))
    ks = adapthist_kernel_size(grey.shape, kernel_size)
    c = _clahe(grey, ks, clip_limit, nbins)
    out = rescale_intensity(c.astype(float_dtype, copy=False), out_range=(0, 1))
    if return_stages:
        return out, {"grey": grey, "clahe": c}
    return out


# ------------------------------------------------------------------ denoise_bilateral
def _gaussian_weight(array, sigma_squared, dtype=float):
    return np.exp(-0.5 * (array ** 2 / sigma_squared), dtype=dtype)


def bilateral_color_lut(bins: int, sigma: float, max_value: float, dtype=float) -> np.ndarray:
    values = np.linspace(0, max_value, bins, endpoint=False)
    return _gaussian_weight(values, sigma ** 2, dtype=dtype)


def bilateral_spatial_lut(win_size: int, sigma: float, dtype=float) -> np.ndarray:
    """ASSUMED: weights on the win_size x win_size window of offsets -(win_size-1)//2 .. +(win_size-1)//2, row-major;
    the exact grid construction of upstream's _compute_spatial_lut is recalled only to this level."""
    ext = (win_size - 1) // 2
    grid = np.arange(-ext, ext + 1)
    rr, cc = np.meshgrid(grid, grid, indexing="ij")
    return _gaussian_weight(np.hypot(rr, cc), sigma ** 2, dtype=dtype).ravel()


def denoise_bilateral(image: np.ndarray, win_size=None, sigma_color=None, sigma_spatial: float = 1, bins: int = 10000,
                      mode: str = "constant", cval: float = 0) -> np.ndarray:
    """skimage.restoration.denoise_bilateral on ONE 2-D single-channel image (channel_axis=None).

    win_size default max(5, 2 ceil(3 sigma_spatial) + 1); sigma_color default image.std(); colour weight looked up in a
    `bins`-entry LUT over [0, max_value) at index min(int(|c - v| * bins / max_value), bins - 1); spatial Gaussian LUT;
    borders by `mode` (default 'constant': outside pixels have value cval and DO take part in the weighted mean);
    images with negative values are shifted by their minimum first and shifted back at the end."""
    image = np.asarray(image)
    if image.ndim != 2:
        raise ValueError("the twin handles one 2-D single-channel image")
    if mode not in ("constant", "edge", "symmetric", "reflect", "wrap"):
        raise ValueError("Invalid mode specified.  Please use `constant`, `edge`, `wrap`, `symmetric` or `reflect`.")
    img = img_as_float(image)
    if img.dtype == np.float16:
        img = img.astype(np.float32)
    fdt = img.dtype.type
    sigma_color = sigma_color or img.std()
    if win_size is None:
        win_size = max(5, 2 * int(math.ceil(3 * sigma_spatial)) + 1)
    min_value, max_value = img.min(), img.max()
    if min_value == max_value:
        return img
    if min_value < 0:
        img = img - min_value
        max_value = max_value - min_value
    color_lut = bilateral_color_lut(bins, sigma_color, max_value, dtype=img.dtype)
    range_lut = bilateral_spatial_lut(win_size, sigma_spatial, dtype=img.dtype)
    ext = (win_size - 1) // 2
    np_mode = {"constant": "constant", "edge": "edge", "symmetric": "symmetric", "reflect": "reflect", "wrap": "wrap"}[mode]
    if np_mode == "constant":
        pad = np.pad(img, ext, mode="constant", constant_values=fdt(cval))
    else:
        pad = np.pad(img, ext, mode=np_mode)
    h, w = img.shape
    dist_scale = fdt(bins / max_value)       # dims == 1
    total_v = np.zeros((h, w), img.dtype)
    total_w = np.zeros((h, w), img.dtype)
    for kr in range(win_size):
        for kc in range(win_size):
            v = pad[kr:kr + h, kc:kc + w]
            t = img - v
            dist = np.sqrt(t * t)
            b = np.minimum((dist * dist_scale).astype(np.int64), bins - 1)
            weight = range_lut[kr * win_size + kc] * color_lut[b]
            total_v += v * weight
            total_w += weight
    out = total_v / total_w
    if min_value < 0:
        out = out + min_value
    return out
