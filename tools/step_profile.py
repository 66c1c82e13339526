#!/usr/bin/env python3
"""Per-kernel device times (deno_profile_begin/end: CUDA events around every launch of libdeno_spectral.so) of one
eager forward + backward of the Darcy FNO2d, warm, no L2 flush.  python tools/step_profile.py"""
import collections
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from deno4pytorch_b200 import _capi  # noqa: E402
from deno4pytorch_b200._lib import lib  # noqa: E402
from deno4pytorch_b200.FNOs import FNO2d  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
B, S, pad = 20, 85, 9
m = FNO2d(1, 1, modes=(12, 12), width=32, depth=4, steps=1, padding=pad).to(dev)
x = torch.randn(B, S, S, 1, device=dev)
g = torch.rand(B, S, S, 2, device=dev)
t = torch.randn(B, S, S, 1, device=dev)
L = lib()
for it in range(4):
    if it == 3:
        L.deno_profile_begin()
    m.zero_grad(set_to_none=True)
    F.mse_loss(m(x, g), t).backward()
    torch.cuda.synchronize()
recs = (_capi.ProfileRecord * 256)()
n = L.deno_profile_end(recs, 256)
agg = collections.OrderedDict()
for r in recs[:min(n, 256)]:
    k = r.name.decode()
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += r.ms * 1e3
tot = sum(v[1] for v in agg.values())
print(f"{n} launches, {tot:.1f} us in libdeno_spectral kernels")
for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{us:9.1f} us  x{c:3d}  {us / c:8.1f} us each  {k}")
