"""Fused enhancement chain: Gaussian denoise -> CLAHE -> unsharp mask
(BASELINE.json config 2; SURVEY.md §7.3 `chain_gcu`).

Equivalent, bit for bit, to
    unsharp_mask(equalize_clahe(gaussian_blur2d(x01, ...), ...), ...)
evaluated on fp32 [0,1] data and quantised once at the end, but executed in two
kernel launches with one byte per pixel of intermediate traffic.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from ._ffi import DTYPE_CODE, as_planes, check, lib, require_cuda, stream_ptr, value_range_of
from .enhance import _check_clahe_args
from .filters import _border, _check_kernel, _pair_float, _pair_int, get_gaussian_kernel1d

__all__ = ["ChainConfig", "ChainPlan", "enhance_chain", "chain_workspace_bytes"]


@dataclass(frozen=True)
class ChainConfig:
    """Keyword names follow the kornia functions each stage replaces."""
    denoise_kernel_size: object = 9          # gaussian_blur2d(kernel_size=)
    denoise_sigma: object = 1.0              # gaussian_blur2d(sigma=)
    clip_limit: float = 2.0                  # equalize_clahe(clip_limit=)
    grid_size: tuple = (8, 8)                # equalize_clahe(grid_size=)
    sharpen_kernel_size: object = 9          # unsharp_mask(kernel_size=)
    sharpen_sigma: object = 1.0              # unsharp_mask(sigma=)
    border_type: str = "reflect"
    value_range: object = None               # (lo, hi) for integer tensors; None = dtype range


def chain_workspace_bytes(n: int, h: int, w: int, grid_size=(8, 8)) -> int:
    return int(lib().mie_chain_workspace_bytes(n, h, w, int(grid_size[0]), int(grid_size[1])))


def enhance_chain(input: torch.Tensor, config: ChainConfig = ChainConfig(), *, out: torch.Tensor = None,
                  out_dtype=None, workspace: torch.Tensor = None, stages: int = 3) -> torch.Tensor:
    """Run the chain on (H,W) / (C,H,W) / (B,C,H,W) planes resident on the GPU.

    `out` and `workspace` (uint8, >= chain_workspace_bytes) may be supplied so that a
    steady-state loop allocates nothing.  `stages` (1 = first launch, 2 = second launch, 3 = both) exists
    for per-kernel timing in bench.py; leave it at 3.
    """
    cfg = config
    _check_clahe_args(cfg.clip_limit, cfg.grid_size)
    gky, gkx = _pair_int(cfg.denoise_kernel_size, "denoise_kernel_size")
    uky, ukx = _pair_int(cfg.sharpen_kernel_size, "sharpen_kernel_size")
    _check_kernel(gky, gkx)
    _check_kernel(uky, ukx)
    gsy, gsx = _pair_float(cfg.denoise_sigma, "denoise_sigma")
    usy, usx = _pair_float(cfg.sharpen_sigma, "sharpen_sigma")
    _border(cfg.border_type)
    require_cuda(input)
    x, n, h, w = as_planes(input)
    lo, hi = value_range_of(x, cfg.value_range)
    if out is None:
        dt = x.dtype if out_dtype is None else out_dtype
        if dt != x.dtype and dt != torch.float32:
            raise TypeError("out_dtype must be the input dtype or torch.float32")
        out = torch.empty(x.shape, dtype=dt, device=x.device)
    elif out.shape != x.shape or not out.is_contiguous() or out.device != x.device or \
            (out.dtype != x.dtype and out.dtype != torch.float32):
        raise ValueError("out must be a contiguous tensor of the input's shape on its device (same dtype or float32)")
    gh, gw = int(cfg.grid_size[0]), int(cfg.grid_size[1])
    L = lib()
    need = L.mie_chain_workspace_bytes(n, h, w, gh, gw)
    if workspace is None:
        workspace = torch.empty(max(need, 1), dtype=torch.uint8, device=x.device)
    elif workspace.dtype != torch.uint8 or workspace.numel() < need or workspace.device != x.device:
        raise ValueError(f"workspace must be a uint8 tensor of >= {need} bytes on the input's device")
    wgx, wgy = get_gaussian_kernel1d(gkx, gsx), get_gaussian_kernel1d(gky, gsy)
    wux, wuy = get_gaussian_kernel1d(ukx, usx), get_gaussian_kernel1d(uky, usy)
    with torch.cuda.device(x.device):
        check(L.mie_chain_gauss_clahe_unsharp(
            x.data_ptr(), out.data_ptr(), DTYPE_CODE[x.dtype], DTYPE_CODE[out.dtype], n, h, w, h * w, w, h * w, w,
            wgx.ctypes.data, gkx, wgy.ctypes.data, gky, gh, gw, float(cfg.clip_limit),
            wux.ctypes.data, ukx, wuy.ctypes.data, uky, _border(cfg.border_type), lo, hi, int(stages),
            workspace.data_ptr(), workspace.numel(), stream_ptr(x.device)))
    return out


class ChainPlan:
    """The chain on FIXED device buffers, captured once into a CUDA graph.

    A steady-state loop (same input / output tensors every step, e.g. the slots of the host pipeline's
    ring) then costs one graph launch per step instead of three kernel launches plus the Python argument
    checks, which matters because the whole config-2 step is ~0.2 ms.  `replay()` enqueues the graph on
    the current stream; `input` is read and `out` is written in place each time.
    """

    def __init__(self, input: torch.Tensor, config: ChainConfig = ChainConfig(), *, out: torch.Tensor = None,
                 out_dtype=None, workspace: torch.Tensor = None, stages: int = 3, graph: bool = True):
        require_cuda(input)
        self.input = input
        self.config = config
        self.stages = int(stages)
        x, n, h, w = as_planes(input)
        if out is None:
            dt = x.dtype if out_dtype is None else out_dtype
            out = torch.empty(input.shape, dtype=dt, device=input.device)
        self.out = out
        gh, gw = int(config.grid_size[0]), int(config.grid_size[1])
        if workspace is None:  # a stage-2-only plan must be given the workspace its stage-1 twin filled
            workspace = torch.empty(max(chain_workspace_bytes(n, h, w, (gh, gw)), 1), dtype=torch.uint8,
                                    device=input.device)
        self.workspace = workspace
        self._graph = None
        self._direct()  # validates the arguments and performs the one-time kernel attribute setup
        if graph and n > 0:
            torch.cuda.synchronize(input.device)
            g = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream(device=input.device)
            with torch.cuda.stream(side):
                with torch.cuda.graph(g, stream=side):
                    self._direct()
            self._graph = g

    def _direct(self):
        enhance_chain(self.input, self.config, out=self.out.view(self.input.shape) if self.out.shape != self.input.shape
                      else self.out, workspace=self.workspace, stages=self.stages)

    def replay(self) -> torch.Tensor:
        if self._graph is not None:
            self._graph.replay()
        else:
            self._direct()
        return self.out
