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
    "pipe2d_w128": (2, dict(in_dim=2, out_dim=1, modes=(12, 12), width=128, depth=4, hidden=128, steps=1, padding=8),
                    (129, 129), 20),
    "turb3d": (3, dict(in_dim=10, out_dim=1, modes=(8, 8, 8), width=32, depth=4, hidden=128, steps=1, padding=6,
                       use_complex=False), (64, 64, 40), 10),
}
WORKLOAD_TEXT = {
    "darcy2d": "Demo/Darcy_2d FNO2d 85x85 (pad 9 -> 94x94), modes (12,12), width 32, depth 4, batch 20 per GPU; "
               "step = fwd + MSE + bwd + Adam",
}


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
    from deno4pytorch_b200.synthetic import synthetic_batch
    ndim, kw, spatial, batch = WORKLOADS[workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    params = fno_port.init_fno_params(ndim, kw["in_dim"], kw["out_dim"], kw["modes"], kw["width"], kw["depth"],
                                      kw["hidden"], kw["steps"], kw.get("use_complex", True), seed=seed)
    trainer = fno_port.PortTrainer(params, ndim=ndim, modes=kw["modes"], depth=kw["depth"], padding=kw["padding"])
    x, grid, target = synthetic_batch(ndim, kw, spatial, batch, seed)
    for _ in range(warmup):
        trainer.step(x, grid, target)
    t0 = time.perf_counter()
    for _ in range(steps):
        trainer.step(x, grid, target)
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps * 1e3, torch.get_num_threads(), batch


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
def layer_roofline(workload: str, device, hbm_peak: float, reps: int = 20):
    """Per-kernel attribution of ONE spectral layer fwd+bwd at the workload's layer shape, measured
    with CUDA events around every launch (deno_profile_begin/end) and, for the layer total, one event
    pair around fwd+bwd.  Algorithmic bytes: DESIGN.md section 4."""
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
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=device)  # > 126 MB L2

    def one():
        for t in (x, lw, lb, *ws):
            t.grad = None
        y = spectral_conv(x, ws, lw, lb, modes, "gelu")
        y.backward(gy)

    for _ in range(3):
        one()
    torch.cuda.synchronize()
    # layer total: event pair around fwd+bwd, L2 flushed before each repetition
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        one()
        e1.record()
        e1.synchronize()
        tot += e0.elapsed_time(e1)
    layer_ms = tot / reps
    # per-kernel attribution
    L = lib()
    acc = {}
    nrec = 64
    recs = (_capi.ProfileRecord * nrec)()
    for _ in range(reps):
        flush.zero_()
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
    algo = {  # algorithmic bytes per launch (read inputs once, write outputs once)
        "plane_analysis": A + spec_small,
        "plane_synthesis": 3 * A + spec_small,              # read x, write y and z
        "plane_analysis_gz": 3 * A + spec_small,            # read gy, z; write gz
        "plane_synthesis_bwd": 2 * A + spec_small,          # read gz, write gx
        "plane_analysis_tc": A + spec_small, "plane_analysis_gz_tc": 3 * A + spec_small,
        "plane_synthesis_tc": 3 * A + spec_small, "plane_synthesis_bwd_tc": 2 * A + spec_small,
        "bypass_wgrad_partial": 2 * A, "bypass_wgrad_tc": 2 * A,   # read gz, x
        "mode_mix_bwd_fused": 3 * spec_small + 3 * Wb,      # read G^, X^, W; write G^x, gW
        "mode_mix": 2 * spec_small + Wb, "mode_mix_bwd": 2 * spec_small + Wb, "mode_mix_wgrad": 2 * spec_small + Wb,
    }
    dom = max(per_kernel, key=per_kernel.get)
    dom_ms = per_kernel[dom] / max(launches[dom], 1)
    dom_bytes = algo.get(dom, 0)
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    layer_bytes = 7 * A + 3 * Wb
    layer_gbs = layer_bytes / (layer_ms * 1e-3) / 1e9
    traffic, traffic_note = None, None
    try:   # dram__bytes_read.sum + dram__bytes_write.sum of the committed `ncu --set full` capture of this kernel
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "ncu_traffic.json")) as f:
            tj = json.load(f)
        if tj.get("workload") == workload and dom in tj["kernels"]:
            traffic = tj["kernels"][dom]["dram_bytes_read"] + tj["kernels"][dom]["dram_bytes_write"]
            traffic_note = tj["source"] + "; " + tj["note"]
    except (OSError, ValueError, KeyError):
        pass
    roof = {
        "bound": "hbm", "kernel": dom, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
        "frac": achieved / hbm_peak, "traffic": traffic, "traffic_source": traffic_note,
        "algorithmic_bytes_per_launch": dom_bytes, "kernel_us": dom_ms * 1e3,
        "layer_fwd_bwd": {"algorithmic_bytes": layer_bytes, "us": layer_ms * 1e3, "achieved": layer_gbs,
                          "frac": layer_gbs / hbm_peak, "formula": "7A+3Wb (SURVEY 8d)"},
        "kernels_us": {k: round(v * 1e3, 2) for k, v in sorted(per_kernel.items(), key=lambda kv: -kv[1])},
        "how": "CUDA events around each launch (deno_profile_begin/end) on the launching stream, L2 flushed "
               "(256 MB memset) before every repetition, mean of %d repetitions" % reps,
    }
    return roof


_OPEN_STEPS = []


def run_own(args):
    from deno4pytorch_b200 import parallel
    from deno4pytorch_b200.functional import launch_count
    from deno4pytorch_b200.training import TrainStep
    import deno4pytorch_b200 as d4
    from deno4pytorch_b200.synthetic import synthetic_batch

    rank, local_rank, world, device = parallel.setup_ddp()
    if device.type != "cuda":
        raise SystemExit("bench.py (own arm) needs a CUDA device; there is no CPU fallback")
    if world != args.gpus and world > 1:
        args.gpus = world
    ndim, kw, spatial, batch = WORKLOADS[args.workload]
    hbm_peak, peak_src = peaks()

    torch.manual_seed(0)  # identical replicas on every rank
    model = {1: d4.FNO1d, 2: d4.FNO2d, 3: d4.FNO3d}[ndim](**kw).to(device)
    nparams = sum(p.numel() * (2 if p.is_complex() else 1) for p in model.parameters())
    x, grid, target = synthetic_batch(ndim, kw, spatial, batch, seed=100 + rank)  # each rank: its own shard
    hx, hg, ht = x.pin_memory(), grid.pin_memory(), target.pin_memory()
    dx, dg, dt_ = x.to(device), grid.to(device), target.to(device)

    before = launch_count()
    step = TrainStep(model, dx, dg, dt_, use_graph=not args.eager, warmup=3)
    _OPEN_STEPS.append(step)
    # kernels of ours per step = launches recorded while one eager step / one capture ran
    probe_before = launch_count()
    if args.eager:
        step()
        per_step_launches = launch_count() - probe_before
    else:
        per_step_launches = (launch_count() - before) // 4  # 3 warm-up steps + 1 captured step
    torch.cuda.synchronize()

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
    # stream while the current step runs, and every step's loss is copied to pinned host memory and READ one step later
    # (read_loss(lag=1)), so the GPU never waits for the host.  Every step still copies its own inputs from pinned
    # host memory and its own loss is read on the host inside the timed region: K copies in, K losses out.
    for _ in range(2):
        float(step(hx, hg, ht))
    step.enable_loss_readback()
    torch.cuda.synchronize()
    barrier()
    step.prefetch(hx, hg, ht)
    t0 = time.perf_counter()
    for k in range(args.steps):
        step()                                   # staged batch -> static buffers -> one optimiser step
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

    if rank != 0:
        if world > 1:
            dist.barrier()
        return

    roof = layer_roofline(args.workload, device, hbm_peak)
    roof["peak_source"] = peak_src
    cpu = None
    if world == 1:
        # bounded CPU sample on rank 0: a few full steps of the same workload (about 10-30 s of CPU work)
        n_cpu = args.cpu_steps
        rate, ms, cores, b = cpu_reference_rate(args.workload, n_cpu, 1)
        cpu = {"value": rate, "unit": "samples/s", "cores": cores, "kind": "port",
               "sample": f"{n_cpu} full train steps at batch {b} after 1 warm-up ({ms:.0f} ms/step), "
                         "oracle/fno_port.py (torch.fft + einsum, the reference's CPU path)"}
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
                   "allreduce": ("per stage, overlapped with the backward, inside the graph" if step._overlap else
                                 ("one flat averaged all-reduce inside the step graph" if getattr(step, "_ingraph", False) else
                                  ("one flat all-reduce between the backward and the optimiser graph" if world > 1 else "none"))),
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
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--workload", default="darcy2d", choices=sorted(WORKLOADS))
    ap.add_argument("--eager", action="store_true", help="no CUDA graph (debug)")
    ap.add_argument("--cpu-steps", type=int, default=8)
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
