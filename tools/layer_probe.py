#!/usr/bin/env python3
"""Runs ONE spectral layer forward+backward a few times at a named config shape (for ncu / sanitizer).
   python tools/layer_probe.py C2 [iters]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from deno4pytorch_b200.functional import spectral_conv  # noqa: E402

SHAPES = {
    "C1": (20, 64, (1026,), (16,)),
    "C2": (20, 32, (94, 94), (12, 12)),
    "C3": (20, 64, (72, 72), (12, 12)),
    "C4a": (20, 128, (229, 59), (12, 12)),
    "C4b": (20, 128, (137, 137), (12, 12)),
    "C5": (10, 32, (70, 70, 46), (8, 8, 8)),
}


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "C2"
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    B, C, spatial, modes = SHAPES[name]
    nd = len(spatial)
    dev = torch.device("cuda", 0)
    gen = torch.Generator().manual_seed(0)
    x = torch.randn((B, C) + spatial, generator=gen).to(dev).requires_grad_(True)
    ws = [torch.view_as_complex(torch.rand((C, C) + modes + (2,), generator=gen) / C ** 2).to(dev).requires_grad_(True)
          for _ in range(2 ** (nd - 1))]
    lw = (torch.randn((C, C) + (1,) * nd, generator=gen) / C ** 0.5).to(dev).requires_grad_(True)
    lb = torch.zeros(C, device=dev, requires_grad=True)
    gy = torch.randn((B, C) + spatial, generator=gen).to(dev)
    for _ in range(iters):
        spectral_conv(x, ws, lw, lb, modes, "gelu").backward(gy)
    torch.cuda.synchronize()
    print("probe ok", name, float(x.grad.abs().mean()))


if __name__ == "__main__":
    main()
