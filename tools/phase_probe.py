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
for it in range(3):
    L.deno_debug_read(buf, 64, 1)
    y = spectral_conv(x, ws, lw, lb, modes, "gelu")
    y.backward(gy)
    torch.cuda.synchronize()
    L.deno_debug_read(buf, 64, 0)
names = ["unit setup", "epi wait D", "epi work", "load wait stage", "load split+sts", "mma wait A_X", "mma wait D free",
         "mma issue", "roles total", "rows"]
vals = list(buf)[:10]
print(name, {n: v for n, v in zip(names, vals)})
rows = max(vals[9], 1)
print("per row cycles:", {n: round(v / rows) for n, v in zip(names[1:9], vals[1:9])}, "setup per unit-row", round(vals[0] / rows))


an = ["wait staged plane", "transform", "cp.async issue", "stage-1 MMA", "D1->B2", "stage-2 MMA", "epilogue", "planes"]
av = list(buf)[32:40]
pl = max(av[7], 1)
print("analysis_tc CTA0 (fwd + gz variants summed):", {n: v for n, v in zip(an, av)})
print("analysis per plane:", {n: round(v / pl) for n, v in zip(an[:7], av[:7])}, "sum", round(sum(av[:7]) / pl))

sn = ["setup", "stage A", "wait prev MMA", "split+sts", "barrier", "mma issue", "prefetch issue", "epilogue(prev row)", "rows"]
sv = list(buf)[48:57]
rw = max(sv[8], 1)
print("synthesis_tc (bulk) CTA0, fwd + bwd-x kernels summed:", {n: v for n, v in zip(sn, sv)})
print("synthesis per row:", {n: round(v / rw) for n, v in zip(sn[2:8], sv[2:8])}, "| per CTA: setup", sv[0] // 2, "stage A", sv[1] // 2)

xv = list(buf)[57:61]
ch = max(xv[3], 1)
print("stage A detail: zero+twiddles per CTA", xv[0] // 2, "| per chunk: load", xv[1] // ch, "compute", xv[2] // ch, "chunks", xv[3])
