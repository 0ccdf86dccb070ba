"""Probe: does write-combined pinned memory (cudaHostAllocWriteCombined) for the UPLOAD staging buffer change
H2D / full-duplex PCIe throughput on this box?  python benchmarks/wc_probe.py"""
import ctypes, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import mie_b200 as M
from mie_b200 import synthetic
from mie_b200.loader import HostSlicePipeline

rt = ctypes.CDLL("libcudart.so.12") if os.path.exists("/usr/local/cuda/lib64/libcudart.so.12") else ctypes.CDLL("libcudart.so")
def host_alloc(nbytes, flags):
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(flags))
    assert rc == 0, rc
    buf = (ctypes.c_char * nbytes).from_address(p.value)
    return torch.frombuffer(buf, dtype=torch.uint8), p

dev = torch.device("cuda:0")
torch.cuda.init()
x_np = synthetic.phantom((256, 1, 512, 512), np.uint16, seed=0)
nbytes = x_np.nbytes
results = {}
for name, flags in (("default", 0), ("write_combined", 4)):
    raw, keep = host_alloc(nbytes, flags)
    x_host = raw.view(torch.uint16).reshape(256, 1, 512, 512)
    x_host.copy_(torch.from_numpy(x_np))
    y_host = torch.empty((256, 1, 512, 512), dtype=torch.uint16).pin_memory()
    print(name, "is_pinned:", x_host.is_pinned())
    xd = torch.empty((256, 1, 512, 512), dtype=torch.uint16, device=dev)
    def t(fn, n=5):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    h2d = t(lambda: xd.copy_(x_host, non_blocking=True))
    pipe = HostSlicePipeline(dev, (512, 512), torch.uint16, chunk=32)
    for _ in range(3): pipe.run(x_host, y_host)
    e2e = t(lambda: pipe.run(x_host, y_host))
    print(name, "h2d ms", round(h2d, 3), "GB/s", round(nbytes / h2d / 1e6, 1), "| e2e ms", round(e2e, 3))
    ref = M.enhance_chain(xd)
    print(name, "e2e output correct:", bool((y_host.to(dev) == ref).all()))
