#!/usr/bin/env python3
"""bench.py - FNO train samples/sec on B200 (BASELINE.json metric) + spectral-conv roofline.

Own arm (default): one "step" = forward + MSE + backward (+ NCCL all-reduce of the flat gradient
buffer under DP) + Adam of the Darcy FNO2d (BASELINE.json configs[1], SURVEY.md C2:
85x85 grid, padding 9, modes (12,12), width 32, depth 4, per-GPU batch 20) on synthetic data.
Weak scaling: every rank trains its own shard of the global batch; no data-path collective.

Reference arm (``--impl reference``): the reference's CPU path (the torch.fft + einsum port in
oracle/fno_port.py, pinned to the real reference by tests/golden) on the same workload, all
host threads, rank 0 only.

Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

WORKLOADS = {
    # name: (ndim, model kwargs, spatial, per-GPU batch)
    "darcy2d": (2, dict(in_dim=1, out_dim=1, modes=(12, 12), width=32, depth=4, hidden=128, steps=1, padding=9),
                (85, 85), 20),
    "burgers1d": (1, dict(in_dim=1, out_dim=1, modes=16, width=64, depth=4, hidden=128, steps=1, padding=2),
                  (1024,), 20),
    "ns2d_w64": (2, dict(in_dim=10, out_dim=1, modes=(12, 12), width=64, depth=4, hidden=128, steps=1, padding=8),
                 (64, 64), 20),
    # BASELINE.json configs[2] as stated: ten autoregressive model calls per training step, one loss, BPTT
    # (Demo/Turbulence_2d+t/run_FNO&UNet.py:61-75); a sample = one 10-step trajectory
    "ns2d_rollout": (2, dict(in_dim=10, out_dim=1, modes=(12, 12), width=64, depth=4, hidden=128, steps=1, padding=8),
                     (64, 64), 20),
    "pipe2d_w128": (2, dict(in_dim=2, out_dim=1, modes=(12, 12), width=128, depth=4, hidden=128, steps=1, padding=8),
                    (129, 129), 20),
    "turb3d": (3, dict(in_dim=10, out_dim=1, modes=(8, 8, 8), width=32, depth=4, hidden=128, steps=1, padding=6,
                       use_complex=False), (64, 64, 40), 10),
}
ROLLOUT_STEPS = {"ns2d_rollout": 10}
WORKLOAD_TEXT = {
    "darcy2d": "Demo/Darcy_2d FNO2d 85x85 (pad 9 -> 94x94), modes (12,12), width 32, depth 4, batch 20 per GPU; "
               "step = fwd + MSE + bwd + Adam",
    "burgers1d": "C1 Demo/Burgers_1d FNO1d grid 1024 (pad 2), modes 16, width 64, batch 20",
    "ns2d_w64": "C3 layer stack, single model call: FNO2d 64x64 (pad 8), modes (12,12), width 64, batch 20",
    "ns2d_rollout": "C3 Demo/Turbulence_2d+t FNO2d 64x64 (pad 8), modes (12,12), width 64, batch 20, TEN autoregressive "
                    "model calls per step (window shift by cat), one MSE over the 10 predictions, BPTT, Adam",
    "pipe2d_w128": "C4b Demo/Pipe_2d FNO2d 129x129 (pad 8 -> 137x137), modes (12,12), width 128, batch 20",
    "turb3d": "C5 Demo/Turbulence_3d+t FNO3d 64x64x40 (pad 6), modes (8,8,8), width 32, real-view weights, batch 10 per GPU",
}
# BASELINE.json config ids of the secondary workloads the default run also measures (N = 1: all; N > 1: C5 only)
SECONDARY = {"C1": "burgers1d", "C3": "ns2d_rollout", "C4b": "pipe2d_w128", "C5": "turb3d"}


def rollout_loss(forward, x, grid, target, T):
    """Demo/Turbulence_2d+t/run_FNO&UNet.py:61-75: T model calls, prediction appended to the window, oldest frame dropped."""
    import torch.nn.functional as F
    xx = x
    preds = []
    for _ in range(T):
        im = forward(xx, grid)
        preds.append(im)
        xx = torch.cat((xx[..., 1:], im), dim=-1)
    return F.mse_loss(torch.cat(preds, dim=-1), target)


def workload_batch(name: str, seed: int):
    from deno4pytorch_b200.synthetic import synthetic_batch
    ndim, kw, spatial, batch = WORKLOADS[name]
    x, grid, target = synthetic_batch(ndim, kw, spatial, batch, seed)
    T = ROLLOUT_STEPS.get(name)
    if T:
        target = torch.randn(target.shape[:-1] + (T * kw["out_dim"],), generator=torch.Generator().manual_seed(seed + 1))
    return x, grid, target


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.thread = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self, extra_steps: int = 0):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        if not self.rows:
            time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons),
                "window": f"timed region + {extra_steps} further untimed steps of the same workload (100 ms nvidia-smi period)"}


# --------------------------------------------------------------------------------------------
# reference arm: the reference's CPU path (oracle port), rank 0 only
# --------------------------------------------------------------------------------------------
def cpu_reference_rate(workload: str, steps: int, warmup: int, seed: int = 0):
    from oracle import fno_port
    ndim, kw, spatial, batch = WORKLOADS[workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    params = fno_port.init_fno_params(ndim, kw["in_dim"], kw["out_dim"], kw["modes"], kw["width"], kw["depth"],
                                      kw["hidden"], kw["steps"], kw.get("use_complex", True), seed=seed)
    trainer = fno_port.PortTrainer(params, ndim=ndim, modes=kw["modes"], depth=kw["depth"], padding=kw["padding"])
    x, grid, target = workload_batch(workload, seed)
    T = ROLLOUT_STEPS.get(workload)

    def one():
        if not T:
            return trainer.step(x, grid, target)
        fwd = lambda a, g: fno_port.fno_forward(trainer.params, a, g, **trainer.cfg)
        loss = rollout_loss(fwd, x, grid, target, T)
        trainer.opt.zero_grad()
        loss.backward()
        trainer.opt.step()
        return loss.item()

    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps * 1e3, torch.get_num_threads(), batch


def torch_gpu_rate(workload: str, device, steps: int = 10, warmup: int = 3, seed: int = 0):
    """The like-for-like GPU baseline: the reference's algorithm on ITS library path - torch.fft (cuFFT) + einsum
    (cuBLAS) + conv (cuDNN) + autograd + torch.optim.Adam - on the same B200, same workload, whole training step, eager
    (how the Demo scripts run it).  oracle/fno_port.py is that algorithm on torch ops; here it is the thing being TIMED
    as a baseline, never a product path."""
    from oracle import fno_port
    ndim, kw, spatial, batch = WORKLOADS[workload]
    params = fno_port.init_fno_params(ndim, kw["in_dim"], kw["out_dim"], kw["modes"], kw["width"], kw["depth"],
                                      kw["hidden"], kw["steps"], kw.get("use_complex", True), seed=seed)
    p = {k: v.to(device).requires_grad_(True) for k, v in params.items()}
    opt = torch.optim.Adam(list(p.values()), lr=1e-3, betas=(0.7, 0.9), weight_decay=1e-4)
    x, grid, target = (t.to(device) for t in workload_batch(workload, seed))
    cfg = dict(ndim=ndim, modes=kw["modes"], depth=kw["depth"], padding=kw["padding"])
    T = ROLLOUT_STEPS.get(workload)
    fwd = lambda a, g: fno_port.fno_forward(p, a, g, **cfg)

    def one():
        loss = rollout_loss(fwd, x, grid, target, T) if T else torch.nn.functional.mse_loss(fwd(x, grid), target)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()

    for _ in range(warmup):
        one()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        one()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    del p, opt
    torch.cuda.empty_cache()
    return {"value": batch / (ms * 1e-3), "unit": "samples/s", "ms_per_step": ms, "steps": steps,
            "what": "reference algorithm on torch ops (cuFFT/cuBLAS/cuDNN + autograd + torch.optim.Adam), eager, same GPU"}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    rate, ms, cores, batch = cpu_reference_rate(args.workload, args.steps, args.warmup)
    sample = f"{args.steps} full train steps of the workload at batch {batch} (after {args.warmup} warm-up)"
    line = {
        "impl": "reference", "metric": "FNO train samples/sec", "value": rate, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_TEXT.get(args.workload, args.workload), "parallelism": "cpu"},
        "cpu_baseline": {"value": rate, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# own arm
# --------------------------------------------------------------------------------------------
ALGO_NOTE = "A = 4*B*C*S bytes (one activation tensor), s = 8*B*C*M (packed spectrum), Wb = 8*C*C*M (complex weights)"


def kernel_algorithmic_bytes(A, spec_small, Wb):
    """Algorithmic bytes per launch of every kernel family (read each input once, write each output once)."""
    return {
        "plane_analysis": A + spec_small, "plane_analysis_tc": A + spec_small, "plane_analysis_tma": A + spec_small,
        "plane_analysis_gz": 3 * A + spec_small, "plane_analysis_gz_tc": 3 * A + spec_small,      # read gy, z'; write gz
        "plane_analysis_gz_tma": 3 * A + spec_small,
        "plane_synthesis": 3 * A + spec_small, "plane_synthesis_tc": 3 * A + spec_small,            # read x, write y and z'
        "plane_synthesis_bwd": 2 * A + spec_small, "plane_synthesis_bwd_tc": 2 * A + spec_small,    # read gz, write gx
        "bypass_wgrad_partial": 2 * A, "bypass_wgrad_tc": 2 * A, "bypass_wgrad_tma": 2 * A,         # read gz, x
        "mode_mix_bwd_fused": 3 * spec_small + 3 * Wb,      # read G^, X^, W; write G^x, gW
        "mode_mix": 2 * spec_small + Wb, "mode_mix_bwd": 2 * spec_small + Wb, "mode_mix_wgrad": 2 * spec_small + Wb,
        "mode_mix_tc": 2 * spec_small + Wb, "mode_mix_bwd_tc": 2 * spec_small + Wb, "mode_mix_wgrad_tc": 2 * spec_small + Wb,
        "mode_mix_lane": 2 * spec_small + Wb, "mode_mix_bwd_lane": 2 * spec_small + Wb,
        "mode_mix_bwd_gx": 2 * spec_small + Wb, "mode_mix_bwd_gw": 2 * spec_small + Wb,   # the fused backward mix split by role (side stream)
        "line_analysis": A + spec_small, "line_analysis_gz": 3 * A + spec_small,
        "row_synthesis": spec_small + A // 2, "row_synthesis_tc": spec_small + A // 2,
    }


def layer_roofline(workload: str, device, hbm_peak: float, reps: int = 20, attribute: bool = True):
    """ONE spectral layer fwd+bwd at the workload's layer shape.
      * ``layer_fwd_bwd``: CUDA-graph replay of the layer (the regime ``value`` runs in: no per-launch events, no host
        launch gaps), one event pair around ``reps`` replays; also the eager, L2-flushed figure for comparison.
      * per-kernel attribution: CUDA events around every launch (deno_profile_begin/end) on the launching stream, warm.
    Algorithmic bytes: DESIGN.md section 4: layer = 7A + 3Wb."""
    from deno4pytorch_b200 import _capi
    from deno4pytorch_b200._lib import lib
    from deno4pytorch_b200.functional import spectral_conv
    ndim, kw, spatial, batch = WORKLOADS[workload]
    C = kw["width"]
    padded = tuple(s + kw["padding"] for s in spatial)
    modes = (kw["modes"],) * ndim if isinstance(kw["modes"], int) else tuple(kw["modes"])
    nblk = 2 ** (ndim - 1)
    S = 1
    for s in padded:
        S *= s
    M = nblk
    for m in modes:
        M *= m
    A = 4 * batch * C * S
    Wb = 8 * C * C * M
    gen = torch.Generator().manual_seed(1)
    x = torch.randn((batch, C) + padded, generator=gen).to(device).requires_grad_(True)
    ws = [torch.view_as_complex(torch.rand((C, C) + modes + (2,), generator=gen) / (C * C)).to(device).requires_grad_(True)
          for _ in range(nblk)]
    lw = (torch.randn((C, C) + (1,) * ndim, generator=gen) / C ** 0.5).to(device).requires_grad_(True)
    lb = torch.randn(C, generator=gen).to(device).requires_grad_(True)
    gy = torch.randn((batch, C) + padded, generator=gen).to(device)

    def one():
        for t in (x, lw, lb, *ws):
            t.grad = None
        y = spectral_conv(x, ws, lw, lb, modes, "gelu")
        y.backward(gy)

    for _ in range(3):
        one()
    torch.cuda.synchronize()
    # (1) graph replay: the regime of the training step
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        one()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        one()
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        graph.replay()
    e1.record()
    e1.synchronize()
    graph_ms = e0.elapsed_time(e1) / reps
    del graph
    # (2) eager with an L2 flush before every repetition (cold-cache figure, includes host launch gaps)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=device)  # > 126 MB L2
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        one()
        e1.record()
        e1.synchronize()
        tot += e0.elapsed_time(e1)
    eager_ms = tot / reps
    del flush
    layer_bytes = 7 * A + 3 * Wb
    layer_gbs = layer_bytes / (graph_ms * 1e-3) / 1e9
    out = {
        "layer_fwd_bwd": {"algorithmic_bytes": layer_bytes, "us": graph_ms * 1e3, "achieved": layer_gbs,
                          "frac": layer_gbs / hbm_peak, "formula": "7A+3Wb (SURVEY 8d)",
                          "regime": "CUDA-graph replay of one layer fwd+bwd, %d replays between one event pair "
                                    "(same regime as `value`)" % reps,
                          "us_eager_l2_flushed": eager_ms * 1e3},
    }
    if not attribute:
        return out
    # (3) per-kernel attribution
    L = lib()
    acc = {}
    nrec = 64
    recs = (_capi.ProfileRecord * nrec)()
    for _ in range(reps):
        torch.cuda.synchronize()
        L.deno_profile_begin()
        one()
        torch.cuda.synchronize()
        n = L.deno_profile_end(recs, nrec)
        for i in range(min(n, nrec)):
            name = recs[i].name.decode()
            acc.setdefault(name, []).append(recs[i].ms)
    per_kernel = {k: sum(v) / reps for k, v in acc.items()}          # ms per layer fwd+bwd
    launches = {k: len(v) // reps for k, v in acc.items()}
    spec_small = 8 * batch * C * M
    algo = kernel_algorithmic_bytes(A, spec_small, Wb)
    dom = max(per_kernel, key=per_kernel.get)
    dom_ms = per_kernel[dom] / max(launches[dom], 1)
    dom_bytes = algo.get(dom, 0)
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    traffic, traffic_note = None, None
    try:   # dram__bytes_read.sum + dram__bytes_write.sum of the committed `ncu --set full` capture of this kernel
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            tj = json.load(f)
        if tj.get("workload") == workload and dom in tj["kernels"]:
            traffic = tj["kernels"][dom]["dram_bytes_read"] + tj["kernels"][dom]["dram_bytes_write"]
            traffic_note = tj["source"] + "; " + tj["note"]
    except (OSError, ValueError, KeyError):
        pass
    out.update({
        "bound": "hbm", "kernel": dom, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
        "frac": achieved / hbm_peak, "traffic": traffic, "traffic_source": traffic_note,
        "algorithmic_bytes_per_launch": dom_bytes, "kernel_us": dom_ms * 1e3, "bytes_note": ALGO_NOTE,
        "kernels_us": {k: round(v * 1e3, 2) for k, v in sorted(per_kernel.items(), key=lambda kv: -kv[1])},
        "kernel_frac": {k: round(algo[k] * launches[k] / (per_kernel[k] * 1e-3) / 1e9 / hbm_peak, 3)
                        for k in per_kernel if k in algo and per_kernel[k] > 0},
        "how": "dominant kernel = largest summed device time in one layer fwd+bwd; CUDA events around each launch "
               "(deno_profile_begin/end) on the launching stream, warm (no flush: one layer's tensors exceed L2), "
               "mean of %d repetitions" % reps,
    })
    return out


_OPEN_STEPS = []


def make_step(workload, device, rank, eager=False):
    """Model + TrainStep of one workload on this rank's shard (weak scaling: the per-GPU batch is the config's)."""
    import deno4pytorch_b200 as d4
    from deno4pytorch_b200.functional import launch_count
    from deno4pytorch_b200.training import TrainStep
    ndim, kw, spatial, batch = WORKLOADS[workload]
    torch.manual_seed(0)  # identical replicas on every rank
    model = {1: d4.FNO1d, 2: d4.FNO2d, 3: d4.FNO3d}[ndim](**kw).to(device)
    nparams = sum(p.numel() * (2 if p.is_complex() else 1) for p in model.parameters())
    x, grid, target = workload_batch(workload, seed=100 + rank)  # each rank: its own shard
    host = (x.pin_memory(), grid.pin_memory(), target.pin_memory())
    T = ROLLOUT_STEPS.get(workload)
    fwd_fn = (lambda m, a, g, t: rollout_loss(m, a, g, t, T)) if T else None
    before = launch_count()
    step = TrainStep(model, x.to(device), grid.to(device), target.to(device), use_graph=not eager, warmup=3,
                     forward_fn=fwd_fn)
    _OPEN_STEPS.append(step)
    if eager:
        b2 = launch_count()
        step()
        per_step = launch_count() - b2
    else:
        per_step = (launch_count() - before) // 4  # 3 warm-up steps + 1 captured step
    torch.cuda.synchronize()
    return step, host, nparams, per_step, batch


def timed_steps(step, nsteps, warmup, world, device):
    """``nsteps`` optimiser steps between one CUDA-event pair, barrier + synchronize on both sides, max over ranks."""
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(nsteps):
        step()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    return ms


def drop_step(step):
    step.close()
    if step in _OPEN_STEPS:
        _OPEN_STEPS.remove(step)
    del step
    import gc
    gc.collect()
    torch.cuda.empty_cache()


def secondary_workload(cfg_id, workload, device, rank, world, hbm_peak, steps=20):
    """One more BASELINE config through the same harness: device-resident samples/s, the layer's graph-replay roofline
    fraction and (rank 0, N = 1) the torch-op baseline on the same GPU."""
    step, host, nparams, per_step, batch = make_step(workload, device, rank)
    ms = timed_steps(step, steps, 3, world, device)
    drop_step(step)
    out = None
    if rank == 0:
        out = {"workload": WORKLOAD_TEXT[workload], "value": world * batch * steps / (ms * 1e-3), "unit": "samples/s",
               "ms_per_step": ms / steps, "steps": steps, "n_gpus": world, "gpu_launches_per_step": per_step,
               "parameters": nparams}
        roof = layer_roofline(workload, device, hbm_peak, reps=10, attribute=True)
        out["layer_fwd_bwd"] = roof["layer_fwd_bwd"]
        out["layer_kernels_us"] = roof.get("kernels_us")
        if world == 1:
            tg = torch_gpu_rate(workload, device, steps=5 if workload != "burgers1d" else 10)
            out["torch_gpu"] = tg
            out["vs_torch_gpu"] = out["value"] / tg["value"]
    torch.cuda.empty_cache()
    return out


def run_own(args):
    from deno4pytorch_b200 import parallel

    rank, local_rank, world, device = parallel.setup_ddp()
    if device.type != "cuda":
        raise SystemExit("bench.py (own arm) needs a CUDA device; there is no CPU fallback")
    if world != args.gpus and world > 1:
        args.gpus = world
    ndim, kw, spatial, batch = WORKLOADS[args.workload]
    hbm_peak, peak_src = peaks()

    step, (hx, hg, ht), nparams, per_step_launches, batch = make_step(args.workload, device, rank, eager=args.eager)

    def barrier():
        if world > 1:
            dist.barrier()

    # ---- device-resident throughput ("value") ----
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    barrier()
    vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
    gpu_index = local_rank
    if vis:
        try:
            gpu_index = int(vis.split(",")[local_rank])
        except (ValueError, IndexError):
            gpu_index = local_rank
    sampler = ClockSampler(gpu_index)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    barrier()
    elapsed_ms = e0.elapsed_time(e1)
    # nvidia-smi reports every 100 ms and the timed region is tens of milliseconds: keep the SAME load running
    # (untimed) until about half a second has been sampled, so the clocks line describes the GPU under this workload
    # and not the idle moment after it.
    extra_steps = 0
    if rank == 0:
        per_step_ms = max(elapsed_ms / max(args.steps, 1), 1e-3)
        extra_steps = int(max(0.0, 500.0 - elapsed_ms) / per_step_ms)
    if world > 1:
        t = torch.tensor([extra_steps], device=device)
        dist.broadcast(t, src=0)
        extra_steps = int(t.item())
    for _ in range(extra_steps):
        step()
    torch.cuda.synchronize()
    clocks = sampler.stop(extra_steps) if rank == 0 else None
    if world > 1:
        t = torch.tensor([elapsed_ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = t.item()
    value = world * batch * args.steps / (elapsed_ms * 1e-3)

    # ---- end to end through the public API: pinned host batch -> H2D -> step -> loss D2H ----
    # Host pipeline of the public API, one step deep: TrainStep.prefetch() starts the NEXT batch's H2D copy on a copy
    # stream while the current step runs (staging buffers; the step then moves them into the graph's static inputs with
    # three small device-to-device copies), and every step's loss is copied to pinned host memory and READ one step
    # later (read_loss(lag=1)), so the GPU never waits for the host.  Every step still copies its own inputs from pinned host memory and its own loss is read on
    # the host inside the timed region: K copies in, K losses out.
    for _ in range(2):
        float(step(hx, hg, ht))
    step.enable_loss_readback()
    torch.cuda.synchronize()
    barrier()
    step.prefetch(hx, hg, ht)
    t0 = time.perf_counter()
    for k in range(args.steps):
        step()                                   # staged batch -> one optimiser step
        step.prefetch(hx, hg, ht)                # H2D of the next step's batch, overlapping this step's kernels
        if k > 0:
            loss_val = step.read_loss(lag=1)     # host value of the previous step's loss
    loss_val = step.read_loss(lag=0)             # ... and of the last one
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = t.item()
    e2e_value = world * batch * args.steps / e2e_s
    h2d = (hx.numel() + hg.numel() + ht.numel()) * 4
    d2h = 4
    allreduce_text = step.allreduce_description()
    drop_step(step)

    # ---- the other BASELINE configs through the same harness (every rank takes part in the DP ones) ----
    others = {}
    if not args.no_secondary and args.workload == "darcy2d":
        ids = list(SECONDARY) if world == 1 else ["C5"]
        for cid in ids:
            try:
                r = secondary_workload(cid, SECONDARY[cid], device, rank, world, hbm_peak)
            except Exception as exc:  # a secondary workload must never take the headline line down with it
                r = {"error": f"{type(exc).__name__}: {exc}"[:300]}
                torch.cuda.empty_cache()
            if rank == 0:
                others[cid] = r

    if rank != 0:
        if world > 1:
            dist.barrier()
        return

    roof = layer_roofline(args.workload, device, hbm_peak)
    roof["peak_source"] = peak_src
    cpu = None
    torch_gpu = None
    if world == 1:
        # bounded CPU sample on rank 0: a few full steps of the same workload (about 10-30 s of CPU work)
        n_cpu = max(1, args.cpu_steps)
        rate, ms, cores, b = cpu_reference_rate(args.workload, n_cpu, 1)
        cpu = {"value": rate, "unit": "samples/s", "cores": cores, "kind": "port",
               "sample": f"{n_cpu} full train steps at batch {b} after 1 warm-up ({ms:.0f} ms/step), "
                         "oracle/fno_port.py (torch.fft + einsum, the reference's CPU path)"}
        torch_gpu = torch_gpu_rate(args.workload, device, steps=20)
    # step working set: saved activations alone exceed L2 -> no explicit flush between steps
    padded = 1
    for s in spatial:
        padded *= s + kw["padding"]
    act_mb = 4 * batch * kw["width"] * padded * (2 * kw["depth"] + 1) / 1e6
    line = {
        "metric": "FNO train samples/sec", "value": value, "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_TEXT.get(args.workload, args.workload), "global_batch": world * batch,
                   "parallelism": f"dp{world}", "parameters": nparams, "cuda_graph": not args.eager,
                   "allreduce": allreduce_text,
                   "l2": f"no explicit flush: saved activations per step ~{act_mb:.0f} MB > 126 MB L2 "
                         "(inputs larger than L2)"},
        "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_s / args.steps * 1e3, "last_loss": loss_val,
                "pipeline": "inputs prefetched one step ahead on a copy stream; each step's loss read on the host one step later"},
        "gpu_launches": per_step_launches * args.steps,
        "gpu_launches_per_step": per_step_launches,
        "clocks": clocks,
        "roofline": roof,
        "cpu_baseline": cpu,
        "torch_gpu": torch_gpu,
        "vs_torch_gpu": (value / torch_gpu["value"]) if torch_gpu else None,
        "workloads": others,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--workload", default="darcy2d", choices=sorted(WORKLOADS))
    ap.add_argument("--eager", action="store_true", help="no CUDA graph (debug)")
    ap.add_argument("--cpu-steps", type=int, default=8)
    ap.add_argument("--no-secondary", action="store_true", help="skip the C1/C3/C4b/C5 lines (quick runs, profiling)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "own":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
        return
    try:
        run_own(args)
    finally:
        for st in _OPEN_STEPS:          # captured graphs may reference NCCL kernels: drop them before the communicator
            st.close()
        if dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
