"""Out-of-bounds WRITE guards for the tuned kernels (compute-sanitizer is not available on the GPU pool):
every output / workspace buffer handed to the C ABI sits inside a larger allocation filled with a sentinel,
and the bytes before and after it must come back untouched.  Ragged and edge shapes on purpose."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GUARD = 4096
SENT = 0xA5


class Guarded:
    """`nbytes` usable bytes at a 256-byte aligned address with GUARD sentinel bytes on both sides."""

    def __init__(self, nbytes, dev):
        self.nbytes = int(nbytes)
        self.buf = torch.full((self.nbytes + 2 * GUARD + 256,), SENT, dtype=torch.uint8, device=dev)
        base = self.buf.data_ptr() + GUARD
        self.off = GUARD + (-base) % 256
        self.ptr = self.buf.data_ptr() + self.off

    def view(self, dtype, shape):
        return self.buf[self.off:self.off + self.nbytes].view(dtype).reshape(shape)

    def check(self, what):
        torch.cuda.synchronize()
        lo = self.buf[:self.off]
        hi = self.buf[self.off + self.nbytes:]
        assert bool((lo == SENT).all()), f"{what}: wrote before the buffer"
        assert bool((hi == SENT).all()), f"{what}: wrote past the buffer"


def _rand(dtype, shape, seed):
    rng = np.random.default_rng(seed)
    if dtype == np.float32:
        return rng.random(shape, dtype=np.float32)
    info = np.iinfo(dtype)
    return rng.integers(info.min, info.max + 1, shape).astype(dtype)


CODE = {np.uint8: 0, np.uint16: 1, np.int16: 2, np.float32: 3}
TORCH = {np.uint8: torch.uint8, np.uint16: torch.uint16, np.int16: torch.int16, np.float32: torch.float32}
RANGE = {np.uint8: (0.0, 255.0), np.uint16: (0.0, 65535.0), np.int16: (-32768.0, 32767.0), np.float32: (0.0, 1.0)}


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


@pytest.mark.parametrize("dtype", [np.uint16, np.int16, np.uint8])
def test_guards_median_equalize_clahe(dev, dtype):
    from mie_b200 import _ffi

    L = _ffi.lib()
    esz = np.dtype(dtype).itemsize
    lo, hi = RANGE[dtype]
    for (n, h, w) in [(3, 40, 48), (2, 33, 264), (1, 2, 8), (5, 64, 512), (2, 17, 1032)]:
        x = torch.from_numpy(_rand(dtype, (n, h, w), 1)).to(dev)
        # 2-D 3x3 median (packed marching kernel for 16-bit), every border rule
        for border in (0, 1, 2, 4):
            for k in (3, 5):
                g = Guarded(n * h * w * esz, dev)
                _ffi.check(L.mie_median2d(x.data_ptr(), g.ptr, CODE[dtype], n, h, w, h * w, w, h * w, w, k, k, border,
                                          _stream(dev)))
                g.check(f"median2d {k}x{k} {(n, h, w)} border {border}")
        # 3x3x3 median on the stack as a volume (direct kernel for even widths), both border rules
        for border in (0, 2):
            g = Guarded(n * h * w * esz, dev)
            _ffi.check(L.mie_median3d(x.data_ptr(), g.ptr, CODE[dtype], n, h, w, h * w, w, h * w, w, None, None,
                                      border, _stream(dev)))
            g.check(f"median3d {(n, h, w)} border {border}")
        # global equalisation: output and per-plane state
        g = Guarded(n * h * w * esz, dev)
        ws = Guarded(L.mie_equalize_workspace_bytes(n), dev)
        _ffi.check(L.mie_equalize(x.data_ptr(), g.ptr, CODE[dtype], CODE[dtype], n, h, w, h * w, w, h * w, w, lo, hi,
                                  ws.ptr, ws.nbytes, _stream(dev)))
        g.check(f"equalize {(n, h, w)}")
        ws.check(f"equalize workspace {(n, h, w)}")
    # CLAHE with enough tiles for the warp-per-tile LUT kernel, and a small job (block-per-tile)
    for (n, h, w, gh, gw) in [(10, 512, 512, 8, 8), (12, 256, 256, 8, 8), (1, 512, 512, 8, 8), (2, 100, 130, 4, 6)]:
        x = torch.from_numpy(_rand(dtype, (n, h, w), 2)).to(dev)
        g = Guarded(n * h * w * esz, dev)
        ws = Guarded(L.mie_clahe_workspace_bytes(n, h, w, gh, gw), dev)
        _ffi.check(L.mie_clahe(x.data_ptr(), g.ptr, CODE[dtype], CODE[dtype], n, h, w, h * w, w, h * w, w, gh, gw, 2.0,
                               0, lo, hi, ws.ptr, ws.nbytes, _stream(dev)))
        g.check(f"clahe {(n, h, w)}")
        ws.check(f"clahe workspace {(n, h, w)}")
        luts = Guarded(n * gh * gw * 256, dev)
        _ffi.check(L.mie_clahe_luts(x.data_ptr(), CODE[dtype], n, h, w, h * w, w, gh, gw, 2.0, 0, lo, hi, luts.ptr,
                                    _stream(dev)))
        luts.check(f"clahe luts {(n, h, w)}")


@pytest.mark.parametrize("dtype", [np.uint16, np.float32])
def test_guards_bilateral_metrics_chain(dev, dtype):
    from mie_b200 import _ffi
    from mie_b200.filters import get_gaussian_kernel1d

    L = _ffi.lib()
    esz = np.dtype(dtype).itemsize
    lo, hi = RANGE[dtype]
    for (n, h, w) in [(2, 70, 50), (1, 33, 65), (1, 128, 256)]:
        x = torch.from_numpy(_rand(dtype, (n, h, w), 3)).to(dev)
        y = torch.from_numpy(_rand(dtype, (n, h, w), 4)).to(dev)
        for k in (3, 9, 15):   # packed kernel (3, 9) and the generic one (15)
            k1 = get_gaussian_kernel1d(k, 1.5)
            wsp = np.ascontiguousarray((k1[:, None] * k1[None, :]).astype(np.float32))
            g = Guarded(n * h * w * 4, dev)
            _ffi.check(L.mie_bilateral(x.data_ptr(), g.ptr, CODE[dtype], 3, n, h, w, h * w, w, h * w, w,
                                       wsp.ctypes.data, k, k, 0.1, 1, lo, hi, _stream(dev)))
            g.check(f"bilateral {(n, h, w)} k={k}")
        out = Guarded(n * 2 * 8, dev)
        ws = Guarded(L.mie_metric_workspace_bytes(n, h, w, 0), dev)
        _ffi.check(L.mie_sqdiff_sums(x.data_ptr(), y.data_ptr(), CODE[dtype], n, h, w, h * w, w, h * w, w, out.ptr,
                                     ws.ptr, ws.nbytes, _stream(dev)))
        out.check("sqdiff out"); ws.check("sqdiff workspace")
        out = Guarded(n * 2 * 8, dev)
        ws = Guarded(L.mie_metric_workspace_bytes(n, h, w, 11), dev)
        _ffi.check(L.mie_ssim_sums(x.data_ptr(), y.data_ptr(), CODE[dtype], n, h, w, h * w, w, h * w, w, 11, 6.5, 58.5,
                                   out.ptr, ws.ptr, ws.nbytes, _stream(dev)))
        out.check("ssim out"); ws.check("ssim workspace")
    # fused chain on the marching path (512-wide) and the tile path (192-wide), output + workspace
    taps = get_gaussian_kernel1d(9, 1.0)
    for (n, h, w, gh, gw) in [(3, 512, 512, 8, 8), (2, 192, 192, 3, 3), (2, 100, 130, 2, 2)]:
        x = torch.from_numpy(_rand(dtype, (n, h, w), 5)).to(dev)
        g = Guarded(n * h * w * esz, dev)
        ws = Guarded(L.mie_chain_workspace_bytes(n, h, w, gh, gw), dev)
        for stages in (3, 3 | 4, 3 | 8):   # default schedule, marching kernels, tile kernels
            g = Guarded(n * h * w * esz, dev)
            ws = Guarded(L.mie_chain_workspace_bytes(n, h, w, gh, gw), dev)
            _ffi.check(L.mie_chain_gauss_clahe_unsharp(x.data_ptr(), g.ptr, CODE[dtype], CODE[dtype], n, h, w, h * w, w,
                                                       h * w, w, taps.ctypes.data, 9, taps.ctypes.data, 9, gh, gw, 2.0,
                                                       taps.ctypes.data, 9, taps.ctypes.data, 9, 1, lo, hi, stages,
                                                       ws.ptr, ws.nbytes, _stream(dev)))
            g.check(f"chain {(n, h, w)} stages {stages}")
            ws.check(f"chain workspace {(n, h, w)} stages {stages}")


def test_guards_clahe16(dev):
    """65 536-bin CLAHE (cluster kernel for small tiles, two-sweep kernel for 256x256-pixel tiles): output and the
    LUT workspace (groups of images) stay inside their buffers."""
    from mie_b200 import _ffi

    L = _ffi.lib()
    for (n, h, w, gh, gw) in [(3, 512, 512, 8, 8), (2, 100, 130, 4, 6), (1, 512, 512, 2, 2), (5, 64, 64, 1, 1)]:
        x = torch.from_numpy(_rand(np.uint16, (n, h, w), 6)).to(dev)
        per_image = L.mie_clahe16_lut_bytes(gh, gw)
        for images_in_ws in (1, 2):
            g = Guarded(n * h * w * 2, dev)
            ws = Guarded(per_image * images_in_ws, dev)
            _ffi.check(L.mie_clahe(x.data_ptr(), g.ptr, 1, 1, n, h, w, h * w, w, h * w, w, gh, gw, 2.0, 1, 0.0, 65535.0,
                                   ws.ptr, ws.nbytes, _stream(dev)))
            g.check(f"clahe16 {(n, h, w)}")
            ws.check(f"clahe16 workspace {(n, h, w)}")
