#!/usr/bin/env python3
"""Runs one forward of a spectral layer at a config shape and prints the phase-timing scratch of the
warp-specialised synthesis kernel (CTA 0).  python tools/phase_probe.py C2"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from deno4pytorch_b200._lib import lib  # noqa: E402
from deno4pytorch_b200.functional import spectral_conv  # noqa: E402
from layer_probe import SHAPES  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
B, Cc, spatial, modes = SHAPES[name]
nd = len(spatial)
dev = torch.device("cuda", 0)
gen = torch.Generator().manual_seed(0)
x = torch.randn((B, Cc) + spatial, generator=gen).to(dev).requires_grad_(True)
ws = [torch.view_as_complex(torch.rand((Cc, Cc) + modes + (2,), generator=gen) / Cc ** 2).to(dev).requires_grad_(True)
      for _ in range(2 ** (nd - 1))]
lw = (torch.randn((Cc, Cc) + (1,) * nd, generator=gen) / Cc ** 0.5).to(dev).requires_grad_(True)
lb = torch.zeros(Cc, device=dev, requires_grad=True)
L = lib()
buf = (C.c_longlong * 64)()
gy = torch.randn((B, Cc) + spatial, generator=gen).to(dev)


def run(which):
    """Phase counters of one forward ("fwd") or one backward ("bwd") after two warm-up passes."""
    for it in range(3):
        y = spectral_conv(x, ws, lw, lb, modes, "gelu")
        torch.cuda.synchronize()
        if which == "fwd" and it == 2:
            L.deno_debug_read(buf, 64, 0)
            return list(buf)
        L.deno_debug_read(buf, 64, 1)
        y.backward(gy)
        torch.cuda.synchronize()
        L.deno_debug_read(buf, 64, 0 if it == 2 else 1)
    return list(buf)


L.deno_debug_read(buf, 64, 1)
for which in ("fwd", "bwd"):
    v = run(which)
    L.deno_debug_read(buf, 64, 1)
    an = ["wait staged plane", "transform", "mma warp idle", "stage-1 wait", "D1->B2", "stage-2 wait", "epilogue", "planes",
          "gz pass", "chunk loop"]
    av = v[32:42]
    pl = max(av[7], 1)
    print(f"[{which}] analysis_tc per plane (CTA 0, {av[7]} planes):",
          {n: round(x_ / pl) for n, x_ in zip(an, av) if n != "planes"})
    sn = ["setup", "stage A", "wait prev MMA", "split+sts", "barrier", "mma issue", "prefetch issue", "epilogue(prev row)", "rows"]
    sv = v[48:57]
    rw = max(sv[8], 1)
    print(f"[{which}] synthesis_tc (bulk) CTA 0: setup {sv[0]} stage A {sv[1]} rows {sv[8]} | per row:",
          {n: round(x_ / rw) for n, x_ in zip(sn[2:8], sv[2:8])})
    xv = v[57:61]
    ch = max(xv[3], 1)
    print(f"[{which}] stage A detail: zero+twiddles {xv[0]} | per chunk: load {xv[1] // ch} compute {xv[2] // ch} chunks {xv[3]}")
    gv = v[10:16]
    if gv[5]:
        print(f"[{which}] bypass_wgrad_tc per tile (CTA 0, {gv[5]} tiles):",
              {n: round(x_ / gv[5]) for n, x_ in zip(["wait free stage", "wait data + split + sts", "addr + load issue",
                                                      "mma wait full", "mma issue"], gv[:5])})
    wv = v[0:10]
    if wv[9]:
        print(f"[{which}] synthesis_ws slots:", wv)


# projection kernels (fno_proj_tc.cuh), Darcy shapes
from deno4pytorch_b200 import _capi  # noqa: E402
from deno4pytorch_b200.FNOs import _ProjFn  # noqa: E402
d = _capi.make_fno_desc(20, 32, 128, 1, 2, 1, (85, 85), 9)
hp = torch.randn(20, 32, 94, 94, device=dev, requires_grad=True)
w1 = (torch.randn(128, 32, device=dev) / 32 ** 0.5).requires_grad_(True)
b1 = torch.zeros(128, device=dev, requires_grad=True)
w2 = (torch.randn(1, 128, device=dev) / 128 ** 0.5).requires_grad_(True)
b2 = torch.zeros(1, device=dev, requires_grad=True)
gyp = torch.randn(20, 85, 85, 1, device=dev)
for it in range(3):
    L.deno_debug_read(buf, 64, 1)
    yp = _ProjFn.apply(hp, w1, b1, w2, b2, d)
    yp.backward(gyp)
    torch.cuda.synchronize()
    L.deno_debug_read(buf, 64, 0)
v = list(buf)
fn = ["wait GEMM1", "A1 store + prefetch", "epilogue math", "exchange + store"]
nt = max(v[20], 1)
print(f"proj_fwd_tc per tile (CTA 0, {v[20]} tiles):", {n: round(x_ / nt) for n, x_ in zip(fn, v[16:20])})
bn = ["wait GEMM1", "chunk math + stores", "ring waits", "A1 store + prefetch", "wait GEMM2", "gh epilogue"]
nt = max(v[27], 1)
print(f"proj_bwd_tc per tile (CTA 0, {v[27]} tiles):", {n: round(x_ / nt) for n, x_ in zip(bn, v[21:27])})
