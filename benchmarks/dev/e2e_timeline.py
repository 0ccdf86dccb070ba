"""Per-chunk timeline of HostSlicePipeline (eager): when does every H2D / kernel group / D2H start and end?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import mie_b200
from mie_b200 import synthetic
from mie_b200.loader import HostSlicePipeline
from mie_b200.chain import enhance_chain
dev = torch.device("cuda:0")
x_host = torch.from_numpy(synthetic.phantom((256, 1, 512, 512), np.uint16, seed=0)).pin_memory()
y_host = torch.empty_like(x_host).pin_memory()
chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 32
pipe = HostSlicePipeline(dev, (512, 512), torch.uint16, chunk=chunk)
for _ in range(3): pipe.run(x_host, y_host, graph=False)
torch.cuda.synchronize()
spans = pipe.schedule(256)
E = lambda: torch.cuda.Event(enable_timing=True)
ev = {k: [(E(), E()) for _ in spans] for k in ("in", "comp", "out")}
t0 = E()
s, d = x_host, y_host
caller = torch.cuda.current_stream(dev)
t0.record(caller)
for st in (pipe.s_in, pipe.s_comp, pipe.s_out): st.wait_stream(caller)
used = [False] * pipe.depth
for i, (z0, z1) in enumerate(spans):
    m, k = z1 - z0, i % pipe.depth
    with torch.cuda.stream(pipe.s_in):
        if used[k]: pipe.s_in.wait_event(pipe.ev_comp[k])
        ev["in"][i][0].record(pipe.s_in)
        pipe.x[k][:m].copy_(s[z0:z1], non_blocking=True)
        ev["in"][i][1].record(pipe.s_in); pipe.ev_in[k].record(pipe.s_in)
    with torch.cuda.stream(pipe.s_comp):
        pipe.s_comp.wait_event(pipe.ev_in[k])
        if used[k]: pipe.s_comp.wait_event(pipe.ev_out[k])
        ev["comp"][i][0].record(pipe.s_comp)
        enhance_chain(pipe.x[k][:m], pipe.config, out=pipe.y[k][:m], workspace=pipe.ws[k])
        ev["comp"][i][1].record(pipe.s_comp); pipe.ev_comp[k].record(pipe.s_comp)
    with torch.cuda.stream(pipe.s_out):
        pipe.s_out.wait_event(pipe.ev_comp[k])
        ev["out"][i][0].record(pipe.s_out)
        d[z0:z1].copy_(pipe.y[k][:m], non_blocking=True)
        ev["out"][i][1].record(pipe.s_out); pipe.ev_out[k].record(pipe.s_out)
    used[k] = True
torch.cuda.synchronize()
for i, (z0, z1) in enumerate(spans):
    row = [f"{z1 - z0:3d}"]
    for k in ("in", "comp", "out"):
        a, b = ev[k][i]
        row.append(f"{k} {t0.elapsed_time(a):6.3f}-{t0.elapsed_time(b):6.3f}")
    print("  ".join(row))
