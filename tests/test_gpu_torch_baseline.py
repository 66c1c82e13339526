"""Times the reference's torch path (cuFFT + cuBLAS + elementwise, via the oracle port) ON THE GPU
next to this repo's kernels - the like-for-like baseline SURVEY.md section 8d asks for.  Lives under
tests/ because only tests may execute oracle/ code; writes gpurun_out/torch_gpu_baseline.json."""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _time(fn, reps=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


@pytest.mark.parametrize("name,B,C,spatial,modes", [
    ("C1", 20, 64, (1026,), (16,)),
    ("C2", 20, 32, (94, 94), (12, 12)),
    ("C3", 20, 64, (72, 72), (12, 12)),
    ("C4b", 20, 128, (137, 137), (12, 12)),
    ("C5", 10, 32, (70, 70, 46), (8, 8, 8)),
])
def test_layer_time_vs_torch_fft_path(name, B, C, spatial, modes):
    from deno4pytorch_b200.functional import spectral_conv
    from oracle import fno_port
    dev = torch.device("cuda", 0)
    nd = len(spatial)
    gen = torch.Generator().manual_seed(0)
    x = torch.randn((B, C) + spatial, generator=gen).to(dev).requires_grad_(True)
    ws = [torch.view_as_complex(torch.rand((C, C) + modes + (2,), generator=gen) / C ** 2).to(dev).requires_grad_(True)
          for _ in range(2 ** (nd - 1))]
    lw = (torch.randn((C, C) + (1,) * nd, generator=gen) / C ** 0.5).to(dev).requires_grad_(True)
    lb = torch.zeros(C, device=dev, requires_grad=True)
    gy = torch.randn((B, C) + spatial, generator=gen).to(dev)

    def ours():
        spectral_conv(x, ws, lw, lb, modes, "gelu").backward(gy)

    def torch_path():
        fno_port.spectral_conv(x, ws, lw, lb, modes, activation="gelu").backward(gy)

    t_ours, t_ref = _time(ours), _time(torch_path)
    S = 1
    for s in spatial:
        S *= s
    M = 2 ** (nd - 1)
    for m in modes:
        M *= m
    algo = 7 * 4 * B * C * S + 3 * 8 * C * C * M
    rec = {"config": name, "ours_us": t_ours * 1e3, "torch_gpu_us": t_ref * 1e3,
           "ours_algo_GBps": algo / t_ours / 1e6, "torch_algo_GBps": algo / t_ref / 1e6, "algo_bytes": algo}
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    path = os.path.join(out, "torch_gpu_baseline.json")
    data = json.load(open(path)) if os.path.exists(path) else {}
    data[name] = rec
    json.dump(data, open(path, "w"), indent=1)
    print(rec)
