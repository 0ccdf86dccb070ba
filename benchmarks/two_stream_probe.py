"""Probe: two independent ChainPlans (own workspace, own output) replayed alternately on two streams, so that the
last partial wave of one step's chain_b overlaps the next step's chain_a (2 048 blocks are 3.46 waves of 592
resident chain_b blocks: 13 % of the SM slots idle in the tail).  python benchmarks/two_stream_probe.py"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import mie_b200 as M
from mie_b200 import synthetic

dev = torch.device("cuda:0")
x = torch.from_numpy(synthetic.phantom((256, 1, 512, 512), np.uint16, 0)).to(dev)
x2 = torch.from_numpy(synthetic.phantom((256, 1, 512, 512), np.uint16, 1)).to(dev)
plans = [M.ChainPlan(x), M.ChainPlan(x2)]
ref = [M.enhance_chain(x), M.enhance_chain(x2)]
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
K = 40
def one_stream():
    for i in range(K): plans[i % 2].replay()
def two_streams():
    cur = torch.cuda.current_stream()
    for s in streams: s.wait_stream(cur)
    for i in range(K):
        with torch.cuda.stream(streams[i % 2]): plans[i % 2].replay()
    for s in streams: cur.wait_stream(s)
for name, fn in (("one stream", one_stream), ("two streams", two_streams)):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    ok = all(torch.equal(p.out, r) for p, r in zip(plans, ref))
    print(name, round(ms, 4), "ms per step", round(256 * 512 * 512 / ms / 1e3, 1), "Mpixel/s", "outputs correct:", ok)
