#!/usr/bin/env python
"""Probe: do the kernels sustain PCIe line rate when they read / write PINNED HOST memory directly (zero copy, UVA)
instead of going through cudaMemcpyAsync staging?  The marching kernels already stream their source rows with bulk
copies (TMA) and store 8 bytes per thread and row, so pointing them at mapped host memory needs no new code.
    python benchmarks/zero_copy_probe.py"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mie_b200 as M  # noqa: E402
from mie_b200 import synthetic  # noqa: E402


class _Alias:
    def __init__(self, t: torch.Tensor):
        self.__cuda_array_interface__ = {"data": (t.data_ptr(), False), "shape": tuple(t.shape),
                                         "typestr": {torch.uint16: "<u2", torch.int16: "<i2", torch.uint8: "|u1",
                                                     torch.float32: "<f4"}[t.dtype], "version": 3, "strides": None}


def device_alias(t: torch.Tensor, dev) -> torch.Tensor:
    """CUDA tensor aliasing a pinned host tensor (unified virtual addressing: same pointer)."""
    assert t.is_pinned()
    return torch.as_tensor(_Alias(t), device=dev)


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    dev = torch.device("cuda:0")
    n = 256
    xh = torch.from_numpy(synthetic.phantom((n, 1, 512, 512), np.uint16, 0)).pin_memory()
    yh = torch.empty_like(xh).pin_memory()
    xd = xh.to(dev); yd = torch.empty_like(xd)
    xa, ya = device_alias(xh, dev), device_alias(yh, dev)
    mb = xh.numel() * 2 / 1e6
    for name, src, dst in (("device -> device", xd, yd), ("HOST -> device (kernel reads pinned host)", xa, yd),
                           ("device -> HOST (kernel writes pinned host)", xd, ya), ("HOST -> HOST", xa, ya)):
        ms = timed(lambda: M.gaussian_blur2d(src, 9, 1.0, out=None) if False else M.unsharp_mask(src, 9, 1.0) if False else None) if False else None
        # gaussian through the C ABI with explicit dst
        from mie_b200.filters import get_gaussian_kernel1d
        from mie_b200._ffi import check, lib, stream_ptr
        w = get_gaussian_kernel1d(9, 1.0)
        def run():
            check(lib().mie_gaussian2d(src.data_ptr(), dst.data_ptr(), 1, 1, n, 512, 512, 512 * 512, 512, 512 * 512, 512,
                                       w.ctypes.data, 9, w.ctypes.data, 9, 1, 0.0, 65535.0, stream_ptr(dev)))
        ms = timed(run)
        print(f"gaussian_blur2d {name:45s} {ms:7.3f} ms  {mb / ms:7.1f} GB/s per direction")
    ref = M.gaussian_blur2d(xd, 9, 1.0)
    print("HOST -> HOST result equals device result:", bool(torch.equal(yh.view(torch.int16), ref.cpu().view(torch.int16))))
    # the chain with host source / host destination, whole batch in one call
    cfg = M.ChainConfig()
    ws = torch.empty(M.chain_workspace_bytes(n, 512, 512), dtype=torch.uint8, device=dev)
    for name, src, dst in (("device -> device", xd, yd), ("HOST -> HOST (zero copy)", xa, ya)):
        ms = timed(lambda: M.enhance_chain(src, cfg, out=dst, workspace=ws))
        print(f"chain {name:30s} {ms:7.3f} ms")
    # chunked, two streams: chain of chunk k+1 overlaps chain of chunk k (reads and writes in both PCIe directions)
    for chunk in (32, 16, 8):
        streams = [torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()]
        wss = [torch.empty(M.chain_workspace_bytes(chunk, 512, 512), dtype=torch.uint8, device=dev) for _ in streams]
        def run_chunks():
            cur = torch.cuda.current_stream()
            for s in streams:
                s.wait_stream(cur)
            for i, z in enumerate(range(0, n, chunk)):
                k = i % len(streams)
                with torch.cuda.stream(streams[k]):
                    M.enhance_chain(xa[z:z + chunk], cfg, out=ya[z:z + chunk], workspace=wss[k])
            for s in streams:
                cur.wait_stream(s)
        run_chunks(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            run_chunks()
        ms = timed(g.replay)
        ok = bool(torch.equal(yh.view(torch.int16), M.enhance_chain(xd, cfg).cpu().view(torch.int16)))
        print(f"chain HOST -> HOST, {chunk:2d}-slice chunks on 3 streams (graph): {ms:7.3f} ms   bit-identical: {ok}")


if __name__ == "__main__":
    main()
