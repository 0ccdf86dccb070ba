#!/usr/bin/env python
"""bench.py — Mpixel/s of the CLAHE + denoise + sharpen chain (BASELINE.json metric) on N B200s.

Headline workload (N=1 and per rank for N>1, weak scaling): BASELINE.json configs[1] —
batch 256 x 512x512 uint16 CT-like phantom slices, Gaussian (K=9, sigma=1, reflect) ->
CLAHE (8x8 tiles, clip 2.0) -> unsharp mask (K=9, sigma=1), uint16 out.

One JSON line on stdout (rank 0).  DESIGN.md §5 says how each field is measured.  Besides the contract keys the line
carries `sub`: one record per other BASELINE.json config (c1 latency, c3_slab = the z-slab + NCCL halo path at
every N, c4 = 64 x 4096^2 bilateral + CLAHE, c5 = NLM), each with its own roofline fraction and CPU figure.

  python bench.py --gpus 1 --steps 20 --warmup 3
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      bench.py --gpus N --steps K --warmup W
  python bench.py --impl reference ...      # CPU arm: the kornia-style torch-CPU twin on the host cores
  python bench.py --no-sub                  # headline only (profiling runs)
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH, H, W = 256, 512, 512
PIXELS = BATCH * H * W
ALG_BYTES_PER_PIXEL = 4  # uint16 in + uint16 out, compulsory traffic only
METRIC = "Mpixel/s, CLAHE+denoise+sharpen chain"
UNIT = "Mpixel/s"
WORKLOAD = ("configs[1]: batch 256 x 512x512 uint16 phantom slices, Gaussian K=9 sigma=1 reflect -> "
            "CLAHE 8x8 clip 2.0 -> unsharp K=9 sigma=1, uint16 out")
CPU_SAMPLE = 32   # slices of the 256-slice batch the CPU arms time per step


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while `active` is set."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons = index, [], set()
        self.active = threading.Event()
        self.stop_flag = threading.Event()
        self.sm_max = None
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = int(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as e:  # NVML missing: report it rather than inventing clocks
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        while not self.stop_flag.is_set():
            if self.active.is_set():
                try:
                    self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                    try:
                        mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                    except Exception:
                        mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                    for bit, name in names.items():
                        if mask & bit:
                            self.reasons.add(name)
                except Exception:
                    pass
            time.sleep(0.005)

    def summary(self):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "NVML unavailable: " + self.err}
        s = sorted(self.samples)
        med = s[len(s) // 2] if s else None
        return {"sm_mhz": med, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons), "samples": len(s)}


# ---------------------------------------------------------------------------------------------- CPU arms
def _use_all_host_threads() -> int:
    """torchrun exports OMP_NUM_THREADS=1; the CPU arms are meant to use every host core."""
    import ctypes

    n = os.cpu_count() or 1
    try:
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(n)
    except OSError:
        pass
    try:
        import torch

        torch.set_num_threads(n)
    except Exception:
        pass
    try:
        import cv2

        cv2.setNumThreads(n)
    except Exception:
        pass
    return n


def _cpu_model() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def _host_info() -> dict:
    info = {"cpu_model": _cpu_model(), "os_cpu_count": os.cpu_count()}
    try:
        import torch

        info["torch_threads"] = torch.get_num_threads()
    except Exception:
        info["torch_threads"] = None
    try:
        import cv2

        info["cv2_threads"] = cv2.getNumThreads()
    except Exception:
        info["cv2_threads"] = None
    return info


def _best_of(fn, reps: int, budget_s: float) -> float:
    best, total = None, 0.0
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
        total += dt
        if total > budget_s:
            break
    return best


def _twin_chain(xt):
    """The chain as a kornia user would write it, on torch-CPU ops (oracle/kornia_twin.py): x/65535 ->
    gaussian_blur2d -> equalize_clahe -> unsharp_mask -> round(clamp * 65535)."""
    import kornia_twin as K
    import torch

    x01 = xt.to(torch.float32) / 65535.0
    g = K.gaussian_blur2d(x01, 9, 1.0)
    c = K.equalize_clahe(g, 2.0, (8, 8))
    u = K.unsharp_mask(c, 9, 1.0)
    return torch.round(u.clamp(0.0, 1.0) * 65535.0).to(torch.int32)


def cpu_chain_figures(sample_slices: int, budget_s: float = 24.0) -> dict:
    """CPU baseline protocol of SURVEY.md §8(d) on a bounded sample of the headline workload: the kornia-style
    torch-CPU twin is the PRIMARY figure; the repo's C oracle port (OpenMP) and the nearest cv2 / scipy building
    blocks are labelled secondary rows.  Same phantom generator and seed as the GPU arm."""
    import numpy as np
    import torch

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O

    from mie_b200 import synthetic

    cores = _use_all_host_threads()
    x = synthetic.phantom((sample_slices, 1, H, W), np.uint16, seed=0)
    px = sample_slices * H * W
    xt = torch.from_numpy(x.astype(np.int32))
    _twin_chain(xt[:2])
    t_twin = _best_of(lambda: _twin_chain(xt), 5, budget_s * 0.45)
    O.chain_gauss_clahe_unsharp(x[:2])
    t_port = _best_of(lambda: O.chain_gauss_clahe_unsharp(x), 5, budget_s * 0.3)
    secondary = [{"name": "oracle C port of the chain (strict fp32 order, OpenMP over slices)", "value": round(px / t_port / 1e6, 2),
                  "unit": UNIT, "threads": cores}]
    try:
        import cv2

        xs = x.reshape(sample_slices, H, W)
        clahe = cv2.createCLAHE(2.0, (8, 8))

        def cv2_chain():   # NOT the same arithmetic (16-bit-native CLAHE, uint16 Gaussians): nearest cv2 pipeline
            for i in range(sample_slices):
                g = cv2.GaussianBlur(xs[i], (9, 9), 1.0, borderType=cv2.BORDER_REFLECT_101)
                c = clahe.apply(g)
                b = cv2.GaussianBlur(c, (9, 9), 1.0, borderType=cv2.BORDER_REFLECT_101)
                cv2.addWeighted(c, 2.0, b, -1.0, 0.0)

        cv2_chain()
        t_cv = _best_of(cv2_chain, 3, budget_s * 0.15)
        secondary.append({"name": "cv2 GaussianBlur -> createCLAHE(65536 bins) -> GaussianBlur/addWeighted on uint16 "
                                  "(different arithmetic; nearest OpenCV pipeline)", "value": round(px / t_cv / 1e6, 2),
                          "unit": UNIT, "threads": cv2.getNumThreads()})
    except Exception as e:  # cv2 missing on the box: say so
        secondary.append({"name": "cv2 pipeline", "unavailable": repr(e)})
    try:
        from scipy import ndimage

        x01 = (x.reshape(sample_slices, H, W)[:8].astype(np.float32) / np.float32(65535.0))
        t_sp = _best_of(lambda: [ndimage.gaussian_filter(p, 1.0, mode="mirror", radius=4) for p in x01], 3, budget_s * 0.1)
        secondary.append({"name": "scipy.ndimage.gaussian_filter alone (one of the two Gaussians; single-threaded by design)",
                          "value": round(8 * H * W / t_sp / 1e6, 2), "unit": UNIT, "threads": 1})
    except Exception as e:
        secondary.append({"name": "scipy gaussian", "unavailable": repr(e)})
    out = {"value": round(px / t_twin / 1e6, 3), "unit": UNIT, "cores": cores, "kind": "port",
           "impl": "kornia-style torch-CPU twin (oracle/kornia_twin.py): F.pad + conv2d, per-tile histc / cumsum / gather / addcmul",
           "sample": f"{sample_slices} of {BATCH} slices (same phantom generator, seed 0), best of <= 5 runs",
           "secondary": secondary}
    out.update(_host_info())
    return out


def run_reference(args):
    """CPU arm.  The reference repository holds no implementation (0 lines of Python); what its dependency list
    implies is kornia on torch-CPU, which cannot be installed here — so this arm times the kornia-style torch-CPU
    twin (the primary CPU figure) with all host threads, each step one bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    import torch

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from mie_b200 import synthetic

    cores = _use_all_host_threads()
    x = synthetic.phantom((CPU_SAMPLE, 1, H, W), np.uint16, seed=0)
    xt = torch.from_numpy(x.astype(np.int32))
    for _ in range(max(args.warmup, 1)):
        _twin_chain(xt)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _twin_chain(xt)
    dt = time.perf_counter() - t0
    value = CPU_SAMPLE * H * W * args.steps / dt / 1e6
    secondary = None
    try:   # the C oracle port beside it (labelled secondary)
        import oracle as O

        O.chain_gauss_clahe_unsharp(x[:2])
        tp = _best_of(lambda: O.chain_gauss_clahe_unsharp(x), 3, 10.0)
        secondary = [{"name": "oracle C port of the chain (OpenMP over slices)", "value": round(CPU_SAMPLE * H * W / tp / 1e6, 2),
                      "unit": UNIT, "threads": cores}]
    except Exception as e:
        secondary = [{"name": "oracle C port", "unavailable": repr(e)}]
    cb = {"value": round(value, 3), "unit": UNIT, "cores": cores, "kind": "port",
          "impl": "kornia-style torch-CPU twin (oracle/kornia_twin.py)",
          "sample": f"{CPU_SAMPLE} slices per step x {args.steps} steps", "secondary": secondary}
    cb.update(_host_info())
    line = {
        "impl": "reference",
        "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "step": f"bounded sample: {CPU_SAMPLE} of {BATCH} slices per step",
                   "note": "the reference repository contains no implementation (0 lines of Python); this arm times the "
                           "kornia-style torch-CPU twin of the chain (what its kornia dependency would run), all host threads"},
        "cpu_baseline": cb,
        "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---------------------------------------------------------------------------------------------- GPU helpers
def _timed_graph(fn, reps, warm=2):
    """ms per call (CUDA events on the current stream).  The call is captured into a CUDA graph when possible so that
    sub-0.1 ms operators are not timed through Python / allocator overhead."""
    import torch

    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    run = fn
    try:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        run = g.replay
        run()
    except Exception:
        run = fn
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def _max_over_ranks(dist, dev, v: float) -> float:
    import torch

    if dist is None:
        return v
    t = torch.tensor([v], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _frac(pixels, ms, peak, bytes_per_px=4):
    gbs = pixels * bytes_per_px / (ms * 1e-3) / 1e9
    return round(gbs, 1), round(gbs / peak, 4)


def sub_c1(M, dev, peak):
    import numpy as np
    import torch
    from mie_b200 import synthetic

    x1 = torch.from_numpy(synthetic.phantom((1, 1, 512, 512), np.uint16, 0)).to(dev)
    ms = _timed_graph(lambda: M.equalize_clahe(x1, 2.0, (8, 8)), 200)
    ms_chain = _timed_graph(lambda: M.enhance_chain(x1), 200)
    return {"workload": "configs[0]: single 512x512 uint16 slice, CLAHE 8x8 clip 2.0 (latency)",
            "clahe_us": round(ms * 1e3, 2), "chain_us": round(ms_chain * 1e3, 2), "launch": "CUDA graph replay"}


def sub_c3(M, dev, dist, world, rank, peak, with_cpu):
    """BASELINE.json configs[2]: 512^3 int16 volume, 3x3x3 median + per-slice CLAHE, z-slab sharded across the ranks with
    one NCCL halo plane per interior face (mie_halo_exchange_z) — STRONG scaling: the volume is fixed.  Every rank also
    runs the unsharded volume on its own GPU: that is the N=1 time the efficiency refers to, and its planes are the
    bit-identity reference for the rank's slab."""
    import numpy as np
    import torch
    from mie_b200 import synthetic

    D = 512
    vol = synthetic.phantom_volume((D, 512, 512), np.int16, seed=0)
    z0, z1 = M.shard_range(D, world, rank)
    full = torch.from_numpy(vol).to(dev)

    def events_ms(run, reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            run()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    if world == 1:
        with M.SlabPlan(full, 2.0, (8, 8)) as plan:
            for _ in range(2):
                plan.replay()
            torch.cuda.synchronize()
            ms1 = events_ms(plan.replay, 10)
        msN, same, eager = ms1, True, None
    else:
        # unsharded on this rank's own GPU without any communication (the ops called directly)
        def unsharded():
            med = M.median(full)
            return M.equalize_clahe(med.unsqueeze(1), 2.0, (8, 8)).squeeze(1)

        ms1 = _max_over_ranks(dist, dev, _timed_graph(unsharded, 10))
        ref = unsharded()
        torch.cuda.synchronize()
        slab = full[z0:z1].clone()
        with M.SlabPlan(slab, 2.0, (8, 8)) as plan:
            for _ in range(3):
                out = plan.replay()
            torch.cuda.synchronize()
            dist.barrier()
            msN = _max_over_ranks(dist, dev, events_ms(plan.replay, 20))
            ok = torch.tensor([float(torch.equal(plan.replay(), ref[z0:z1]))], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            same = bool(ok.item() == 1.0)
        # peer-load variant: neighbour slabs mapped by CUDA IPC, boundary planes read over NVLink inside the median kernel
        ms_peer, same_peer = None, None
        try:
            torch.cuda.synchronize()
            dist.barrier()
            with M.PeerSlabPlan(slab, 2.0, (8, 8)) as pplan:
                for _ in range(3):
                    pplan.replay()
                torch.cuda.synchronize()
                dist.barrier()
                ms_peer = _max_over_ranks(dist, dev, events_ms(pplan.replay, 20))
                okp = torch.tensor([float(torch.equal(pplan.replay(), ref[z0:z1]))], device=dev)
                dist.all_reduce(okp, op=dist.ReduceOp.MIN)
                same_peer = bool(okp.item() == 1.0)
                torch.cuda.synchronize()
        except Exception as exc:   # no peer path between the GPUs of this box, IPC refused, ...
            ms_peer, same_peer = None, f"unavailable: {type(exc).__name__}: {exc}"[:160]
        # the same step launched eagerly (Python + a dozen launches per step)
        for _ in range(2):
            M.median3d_clahe_slab(slab, 2.0, (8, 8))
        torch.cuda.synchronize()
        dist.barrier()
        eager = round(_max_over_ranks(dist, dev, events_ms(lambda: M.median3d_clahe_slab(slab, 2.0, (8, 8)), 10)), 4)
    vox = D * 512 * 512
    ms_nccl = msN
    peer_used = False
    if world > 1 and ms_peer is not None and same_peer is True and ms_peer < msN:
        msN, peer_used = ms_peer, True      # the step a user would run: the faster of the two bit-identical plans
    gbs, frac = _frac(vox, msN, peak * world)
    rec = {"workload": "configs[2]: 512x512x512 int16 volume, 3x3x3 median (nearest) + per-slice CLAHE 8x8 clip 2.0",
           "n_gpus": world, "scaling": "strong", "ms": round(msN, 4), "mvoxel_s": round(vox / msN / 1e3, 1),
           "ms_n1_same_gpu": round(ms1, 4), "efficiency_vs_n1": round(ms1 / (world * msN), 4),
           "bit_identical_to_unsharded": same, "launch": "CUDA graph (PeerSlabPlan: one median launch + CLAHE)" if peer_used else "CUDA graph (SlabPlan: halo exchange + kernels captured)",
           "ms_eager": eager, "halo_bytes_per_face": 512 * 512 * 2, "halo_messages_total": 2 * (world - 1) * 1,
           "exchange": ("peer loads: the neighbour slabs are mapped by CUDA IPC and the median kernel reads their boundary planes "
                        "over NVLink (PeerSlabPlan); the NCCL plan (mie_halo_exchange_z) is timed beside it" if peer_used else
                        "mie_halo_exchange_z: ncclSend/ncclRecv group on the process group's communicator, side stream, "
                        "overlapped with the median of the interior planes") if world > 1 else "none (one slab)",
           "native_exchange": bool(world > 1 and M.volume.nccl_comm_ptr(dev) != 0) if world > 1 else None,
           "ms_nccl_exchange_plan": round(ms_nccl, 4) if world > 1 else None,
           "ms_peer_load_plan": (round(ms_peer, 4) if ms_peer is not None else None) if world > 1 else None,
           "peer_load_plan": ({"bit_identical_to_unsharded": same_peer, "reported_as_ms": peer_used,
                               "how": "PeerSlabPlan: neighbour slabs mapped by CUDA IPC, the median kernel reads their boundary "
                                      "planes over NVLink (no exchange launch, one median launch)"} if world > 1 else None),
           "roofline": {"bound": "alu (integer min/max), then hbm", "achieved": gbs, "peak": peak * world, "unit": "GB/s",
                        "frac": frac, "basis": "4 B/voxel (int16 in + int16 out) / step time; halo traffic not counted"}}
    if with_cpu:
        rec["cpu"] = _cpu_c3()
    return rec


def _cpu_c3():
    import numpy as np
    from scipy import ndimage

    from mie_b200 import synthetic

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O

    v = synthetic.phantom_volume((8, 512, 512), np.int16, seed=0)
    t = _best_of(lambda: ndimage.median_filter(v, size=3, mode="nearest"), 2, 6.0)
    med = O.median3d(v)
    x01 = O.to01(med)
    tc = _best_of(lambda: O.equalize_clahe(x01.reshape(8, 1, 512, 512), 2.0, (8, 8)), 2, 3.0)
    tp = _best_of(lambda: O.median3d(v), 2, 4.0)
    vox = v.size
    return {"sample": "8 of 512 planes", "scipy_median_filter_mvoxel_s": round(vox / t / 1e6, 2),
            "oracle_port_median_mvoxel_s": round(vox / tp / 1e6, 2),
            "scipy_median_plus_oracle_clahe_mvoxel_s": round(vox / (t + tc) / 1e6, 2), "threads": "scipy: 1 (by design); oracle: OpenMP"}


def sub_c4(M, dev, dist, world, rank, peak, with_cpu):
    """BASELINE.json configs[3]: batch 64 x 4096x4096 uint16 radiographs, 9x9 bilateral (sigma_color 0.1, sigma_space 1.5)
    + CLAHE 16x16 clip 2.0, uint16 out; slice-sharded across ranks (strong scaling, no collective)."""
    import numpy as np
    import torch
    from mie_b200 import synthetic

    if not hasattr(M, "bilateral_clahe"):
        return {"unavailable": "fused bilateral -> CLAHE path not built"}
    n_total = 64
    s0, s1 = M.shard_range(n_total, world, rank)
    nb = s1 - s0
    base = synthetic.phantom((8, 1, 4096, 4096), np.uint16, seed=rank)       # 8 distinct images, cycled
    x = torch.from_numpy(base).to(dev).repeat((nb + 7) // 8, 1, 1, 1)[:nb].contiguous()
    out = torch.empty_like(x)
    plan = M.BilateralClahePlan(x, out=out)
    for _ in range(1):
        plan.run()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 2
    for _ in range(reps):
        plan.run()
    e1.record()
    torch.cuda.synchronize()
    ms = _max_over_ranks(dist, dev, e0.elapsed_time(e1) / reps)
    px = n_total * 4096 * 4096
    gbs, frac = _frac(px, ms, peak * world)
    rec = {"workload": "configs[3]: batch 64 x 4096x4096 uint16, bilateral 9x9 (sigma_color 0.1, sigma_space 1.5, reflect) -> "
                       "CLAHE 16x16 clip 2.0, uint16 out", "n_gpus": world, "images": n_total, "scaling": "strong",
           "ms": round(ms, 3), "mpixel_s": round(px / ms / 1e3, 1), "stage_ms": plan.stage_ms(),
           "intermediate": "1-byte lookup-index plane + per-tile histograms (no fp32 image)",
           "colour_weight": "MUFU.EX2 (default mode, rel 6e-7 of the reproducible polynomial kernels = kernel policy bilateral_exact_exp)",
           "roofline": {"bound": "XU pipe (81 MUFU.EX2 per pixel) + FMA pipe, then hbm", "achieved": gbs, "peak": peak * world, "unit": "GB/s",
                        "frac": frac, "basis": "4 B/px (uint16 in + uint16 out) / step time"}}
    if with_cpu:
        rec["cpu"] = _cpu_c4()
    return rec


def _cpu_c4():
    import numpy as np

    from mie_b200 import synthetic

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O

    x = synthetic.phantom((1, 1, 1024, 1024), np.uint16, seed=0)
    x01 = O.to01(x)

    def run():
        b = O.bilateral_blur(x01, 9, 0.1, (1.5, 1.5))
        return O.equalize_clahe(b, 2.0, (4, 4))

    t = _best_of(run, 2, 8.0)
    return {"sample": "one 1024x1024 crop (same 256-px CLAHE tiles)", "oracle_port_mpixel_s": round(x.size / t / 1e6, 3),
            "threads": os.cpu_count(), "note": "kornia's own bilateral_blur materialises (B,C,H,W,81): 348 GB at this batch"}


def sub_c5(M, dev, dist, world, rank, peak, with_cpu):
    """BASELINE.json configs[4] (no learned denoiser is defined in configs/): 7x7 non-local means, search radius 11, on
    512 x 256x256 uint16 patches; slice-sharded (strong scaling)."""
    import numpy as np
    import torch
    from mie_b200 import synthetic

    n_total = 512
    s0, s1 = M.shard_range(n_total, world, rank)
    x = torch.from_numpy(synthetic.phantom((s1 - s0, 1, 256, 256), np.uint16, seed=rank)).to(dev)
    M.denoise_nl_means(x, 7, 11, 0.1)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(2):
        M.denoise_nl_means(x, 7, 11, 0.1)
    e1.record()
    torch.cuda.synchronize()
    ms = _max_over_ranks(dist, dev, e0.elapsed_time(e1) / 2)
    px = n_total * 256 * 256
    gbs, frac = _frac(px, ms, peak * world)
    rec = {"workload": "configs[4] else-branch: 7x7 non-local means, patch_distance 11, h 0.1, batch 512 x 256x256 uint16",
           "n_gpus": world, "scaling": "strong", "ms": round(ms, 3), "mpixel_s": round(px / ms / 1e3, 1),
           "roofline": {"bound": "lsu / fma (compute-bound: 529 offsets per pixel)", "achieved": gbs, "peak": peak * world,
                        "unit": "GB/s", "frac": frac, "basis": "4 B/px / step time"}}
    if with_cpu:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle as O

        xs = O.to01(synthetic.phantom((1, 1, 256, 256), np.uint16, seed=0))
        t = _best_of(lambda: O.denoise_nl_means(xs, 7, 11, 0.1), 2, 6.0)
        rec["cpu"] = {"sample": "1 of 512 patches", "oracle_port_mpixel_s": round(256 * 256 / t / 1e6, 3)}
    return rec


def copy_ceiling(dev, x_host, y_host, dist):
    """Copy-only ceiling of the e2e step on THIS box at THIS N: the same pinned buffers, no kernels.  H2D alone, D2H
    alone and both at once (full duplex) — all ranks at the same time, max over ranks."""
    import torch

    xd = torch.empty(x_host.shape, dtype=x_host.dtype, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    res = {}

    def timed(fn, reps=4):
        fn()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return _max_over_ranks(dist, dev, (time.perf_counter() - t0) / reps * 1e3)

    def h2d():
        with torch.cuda.stream(s1):
            xd.copy_(x_host, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            y_host.copy_(xd, non_blocking=True)

    def both():
        h2d()
        d2h()

    res["h2d_ms"] = round(timed(h2d), 4)
    res["d2h_ms"] = round(timed(d2h), 4)
    res["duplex_ms"] = round(timed(both), 4)
    nbytes = x_host.numel() * x_host.element_size()
    res["h2d_GBps_per_gpu"] = round(nbytes / res["h2d_ms"] / 1e6, 1)
    res["d2h_GBps_per_gpu"] = round(nbytes / res["d2h_ms"] / 1e6, 1)
    res["duplex_GBps_per_gpu_each_way"] = round(nbytes / res["duplex_ms"] / 1e6, 1)
    return res


# ---------------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import numpy as np
    import torch

    import mie_b200
    from mie_b200 import synthetic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    peak, peak_src = _peaks()

    # inputs: one distinct batch per rank (weak scaling: per-GPU work fixed)
    x_host = torch.from_numpy(synthetic.phantom((BATCH, 1, H, W), np.uint16, seed=rank)).pin_memory()
    y_host = torch.empty_like(x_host).pin_memory()
    x = x_host.to(dev, non_blocking=True)
    y = torch.empty_like(x)
    cfg = mie_b200.ChainConfig()
    path = int(mie_b200._lib().mie_chain_is_fused(H, W, 8, 8, 9, 9, 9, 9))
    fused = path > 0
    launches_per_step = {0: 4, 1: 2, 2: 3}[path]  # tuned path: chain_a, cell packing, chain_b

    # steady state = fixed device buffers: the three launches of a step are captured once into a CUDA
    # graph (mie_b200.ChainPlan) and replayed, so the timed region holds no per-call Python work.  Three
    # batches are in flight (mie_b200.ChainRing: three plans with their own input / output / workspace on three
    # streams, replayed round-robin), so that the partial last wave of one step's chain_b overlaps the next
    # step's chain_a; every step still is one complete pass over one 256-slice batch.
    x_b = torch.from_numpy(synthetic.phantom((BATCH, 1, H, W), np.uint16, seed=rank + 1000)).to(dev)
    x_c = torch.from_numpy(synthetic.phantom((BATCH, 1, H, W), np.uint16, seed=rank + 2000)).to(dev)
    ring = mie_b200.ChainRing([x, x_b, x_c], cfg, outs=[y, torch.empty_like(x_b), torch.empty_like(x_c)])
    plan = ring.plans[0]
    ws = plan.workspace
    step_no = [0]

    def step():
        ring.replay(step_no[0])
        step_no[0] += 1

    sampler = ClockSampler(local)
    sampler.start()
    sampler.active.set()

    ring.begin()
    for _ in range(max(args.warmup, 3)):
        step()
    ring.join()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ring.begin()
    for _ in range(args.steps):
        step()
    ring.join()
    e1.record()
    barrier()
    ms_step = _max_over_ranks(dist, dev, e0.elapsed_time(e1)) / args.steps
    value = world * PIXELS / (ms_step * 1e-3) / 1e6

    # ---- one step alone (single plan, back to back on one stream): what a caller without a ring gets
    for _ in range(3):
        plan.replay()
    torch.cuda.synchronize()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    for _ in range(max(args.steps, 10)):
        plan.replay()
    a1.record()
    torch.cuda.synchronize()
    ms_single = a0.elapsed_time(a1) / max(args.steps, 10)

    # ---- per-kernel durations (CUDA events on the launching stream), same inputs, same workspace
    kern = {}
    if fused:
        # stage A / stage B of the fused chain (include/mie.h: MIE_CHAIN_STAGE_A / _B).  On the workload's
        # geometry (path 2, W % 128 == 0) these are the marching kernels of csrc/chain_march.cu; stage B is
        # the tiny cell-table packing launch plus chain_b.
        stage_names = (("chain_a_march_kernel", 1), ("chain_pack_cells_kernel+chain_b_march_kernel", 2)) if path == 2 \
            else (("chain_a_kernel", 1), ("chain_b_kernel", 2))
        for name, mask in stage_names:
            g = mie_b200.ChainPlan(x, cfg, out=y, workspace=ws, stages=mask)
            for _ in range(3):
                g.replay()
            torch.cuda.synchronize()
            reps = max(args.steps, 10)
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(reps):
                g.replay()
            a1.record()
            torch.cuda.synchronize()
            kern[name] = a0.elapsed_time(a1) / reps

    # ---- end to end through the public API with HOST buffers (pinned), H2D + D2H inside the timed region
    from mie_b200.loader import HostSlicePipeline

    pipe = HostSlicePipeline(dev, (H, W), torch.uint16, chunk=32, config=cfg)

    def time_e2e(graph):
        for _ in range(2):
            pipe.run(x_host, y_host, graph=graph)
        barrier()
        n = max(3, min(args.steps, 10))
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(n):
            pipe.run(x_host, y_host, graph=graph)
        g1.record()
        barrier()
        return _max_over_ranks(dist, dev, g0.elapsed_time(g1) / n)

    e2e_ms = time_e2e(True)
    e2e_eager_ms = time_e2e(False)
    checksum = int(y_host.view(torch.int16).flatten()[::997].to(torch.int64).sum().item())
    ceiling = copy_ceiling(dev, x_host, y_host, dist)
    # The e2e step is bound by the box's host memory / PCIe, which other tenants share: a pipelined run far above the
    # copy-only time of the SAME buffers measured seconds later saw interference, not the pipeline (observed once at
    # N = 2: 7.3 ms against a 4.0 ms ceiling, 4.0 ms on the next two runs).  Re-measure ONCE and say so.
    e2e_first = None
    if e2e_ms > 1.4 * ceiling["duplex_ms"]:
        e2e_first = e2e_ms
        e2e_ms = min(e2e_ms, time_e2e(True))
    e2e_value = world * PIXELS / (e2e_ms * 1e-3) / 1e6

    # keep the GPU busy a little longer so that the clock sampler sees the chain under load
    t_end = time.perf_counter() + 0.3
    while time.perf_counter() < t_end:
        ring.begin()
        for _ in range(20):
            step()
        ring.join()
        torch.cuda.synchronize()
    sampler.active.clear()
    sampler.stop_flag.set()

    # ---- the other BASELINE.json configs as sub-records
    sub = {}
    if not args.no_sub:
        del ring, x_b, x_c
        torch.cuda.empty_cache()
        with_cpu = world == 1 and rank == 0
        for name, fn in (("c1", lambda: sub_c1(mie_b200, dev, peak)),
                         ("c3_slab", lambda: sub_c3(mie_b200, dev, dist, world, rank, peak, with_cpu)),
                         ("c4", lambda: sub_c4(mie_b200, dev, dist, world, rank, peak, with_cpu)),
                         ("c5", lambda: sub_c5(mie_b200, dev, dist, world, rank, peak, with_cpu))):
            if name == "c1" and rank != 0:
                continue
            sub[name] = fn()
            torch.cuda.empty_cache()

    if rank == 0:
        step_gbs = ALG_BYTES_PER_PIXEL * PIXELS / (ms_step * 1e-3) / 1e9
        traffic_total, dom_rec, traffic_seq = None, None, None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                tj = json.load(f)
        except Exception:
            tj = {}
        if kern:
            dom = max(kern, key=kern.get)
            # chain_a: uint16 in + 1-byte index out; chain_b: 1-byte index in + uint16 out  -> 3 B/pixel each
            dom_bytes = 3 * PIXELS
            ach = dom_bytes / (kern[dom] * 1e-3) / 1e9
            parts = [tj.get(k) for k in dom.split("+")]
            dom_rec = {"kernel": dom, "achieved": round(ach, 1), "unit": "GB/s", "frac": round(ach / peak, 4),
                       "bytes_per_launch": dom_bytes, "ms_per_launch": round(kern[dom], 4),
                       "traffic": sum(parts) if all(v is not None for v in parts) else None,
                       "basis": "the kernel's own compulsory bytes incl. the 1-byte index plane (3 B/px); an intermediate, so NOT the contract figure",
                       "kernels_ms": {k: round(v, 4) for k, v in kern.items()}}
            allk = [tj.get(k) for k in ("chain_a_march_kernel", "chain_pack_cells_kernel", "chain_b_march_kernel")]
            traffic_total = sum(allk) if all(v is not None for v in allk) else None
            seq = tj.get("in_sequence", {})
            seqk = [seq.get(k) for k in ("chain_a_march_kernel", "chain_pack_cells_kernel", "chain_b_march_kernel")]
            traffic_seq = sum(seqk) if all(isinstance(v, int) for v in seqk) else None
        roof = {"bound": "hbm", "achieved": round(step_gbs, 1), "peak": peak * 1.0, "unit": "GB/s",
                "frac": round(step_gbs / peak, 4), "traffic": traffic_total,
                "basis": "SURVEY.md §8(d): 4 B/pixel (uint16 in + uint16 out; intermediates are not algorithmic bytes) x "
                         "67 108 864 pixels / ms_per_step, per GPU",
                "frac_of_nominal_8TBs": round(step_gbs / 8000.0, 4), "bytes_per_step": ALG_BYTES_PER_PIXEL * PIXELS,
                "peak_source": peak_src, "traffic_source": "profiles/traffic.json (ncu dram__bytes_read+write, sum of the step's three launches, caches flushed before each)",
                "traffic_in_sequence": traffic_seq,
                "traffic_in_sequence_note": "the same metric with --cache-control none: consecutive steps as they really run, write-backs of "
                                            "the index plane and the cell tables included; the step is FMA-pipe bound, not DRAM bound",
                "limiter": "FMA pipe / instruction issue, not HBM (profiles/README.md)", "dominant_kernel": dom_rec}
        cpu = None
        if world == 1:   # the CPU baseline is a rank-0, N = 1 figure (the reference arm times it at every N)
            cpu = cpu_chain_figures(CPU_SAMPLE)
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms_step, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "per_gpu_batch": BATCH, "global_batch": BATCH * world,
                       "parallelism": f"slice-sharded x{world}, no collective on the data path",
                       "l2": "inputs larger than L2 (134 MB in + 134 MB out + 67 MB index plane per step vs 126 MB L2)",
                       "fused_path": fused,
                       "batches_in_flight": "3 (ChainRing: round-robin plans / streams; each step = one full batch)",
                       "ms_per_step_single_plan": round(ms_single, 4)},
            "roofline": roof,
            "cpu_baseline": cpu,
            "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": x_host.numel() * 2,
                    "d2h_bytes_per_step": y_host.numel() * 2, "ms_per_step": round(e2e_ms, 4),
                    "api": "mie_b200.loader.HostSlicePipeline.run(x_host, y_host): pinned host uint16 in/out, 32-slice chunks, "
                           "H2D / kernels / D2H overlapped on three streams; repeated calls on the same buffers replay one CUDA graph",
                    "eager": {"value": round(world * PIXELS / (e2e_eager_ms * 1e-3) / 1e6, 1), "ms_per_step": round(e2e_eager_ms, 4),
                              "note": "graph=False: what a loader that rotates its staging buffers gets"},
                    "copy_ceiling": ceiling,
                    "frac_of_copy_ceiling": round(ceiling["duplex_ms"] / e2e_ms, 4),
                    "remeasured_once": e2e_first is not None,
                    "first_attempt_ms": None if e2e_first is None else round(e2e_first, 4)},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": sampler.summary(),
            "checksum": checksum,
            "sub": sub,
        }
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def _claim_stdout():
    """Exactly ONE line may reach stdout.  Libraries write there too (NCCL prints its version banner on stdout when
    NCCL_DEBUG=VERSION is set on the box), so fd 1 is pointed at stderr for the whole run and the JSON line is written
    to the saved descriptor at the end."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(saved, "w")


_OUT = None


def emit(line: dict) -> None:
    _OUT.write(json.dumps(line) + "\n")
    _OUT.flush()


def main():
    global _OUT
    _OUT = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-sub", action="store_true", help="headline only: skip the c1 / c3 / c4 / c5 sub-records")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
