#!/usr/bin/env python3
"""Times the non-spectral parts of one Darcy training step (lift + pad, crop + projection, loss, optimiser)
with CUDA events, warm, as they run today (torch ops + _TallLinearFn).  python tools/liftproj_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from deno4pytorch_b200.FNOs import FNO2d, _tall_linear  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
B, S, pad, width = 20, 85, 9, 32
m = FNO2d(1, 1, modes=(12, 12), width=width, depth=4, steps=1, padding=pad).to(dev)
x = torch.randn(B, S, S, 1, device=dev)
g = torch.rand(B, S, S, 2, device=dev)


def timeit(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def lift():
    h = _tall_linear(torch.cat((x, g), dim=-1), m.fc0).movedim(-1, 1)
    return F.pad(h, [0, pad, 0, pad])


hp = lift().detach()
gh = torch.randn_like(hp)


def lift_fb():
    for p in m.fc0.parameters():
        p.grad = None
    lift().backward(gh)


hin = torch.randn(B, width, S + pad, S + pad, device=dev, requires_grad=True)


def proj():
    h = hin[..., :-pad, :-pad].movedim(1, -1)
    return _tall_linear(F.gelu(_tall_linear(h, m.fc1)), m.fc2)


gy = torch.randn(B, S, S, 1, device=dev)


def proj_fb():
    hin.grad = None
    for p in list(m.fc1.parameters()) + list(m.fc2.parameters()):
        p.grad = None
    proj().backward(gy)


print("lift fwd us", round(timeit(lambda: lift()), 1), "| lift fwd+bwd us", round(timeit(lift_fb), 1))
print("proj fwd us", round(timeit(lambda: proj()), 1), "| proj fwd+bwd us", round(timeit(proj_fb), 1))
