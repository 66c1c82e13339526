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


def _record(name, rec):
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    path = os.path.join(out, "torch_gpu_baseline.json")
    data = json.load(open(path)) if os.path.exists(path) else {}
    data[name] = rec
    json.dump(data, open(path, "w"), indent=1)
    print(rec)


@pytest.mark.parametrize("name,shape,nb,act", [
    ("afno1d", (10, 64, 128), 4, "relu"),            # the reference's own smoke shapes (spectral_layers.py:599-606)
    ("afno2d", (10, 64, 55, 64), 4, "gelu"),
])
def test_afno_time_vs_torch_fft_path(name, shape, nb, act):
    """Adaptive Fourier layer fwd + bwd: deno_afno_fwd/_bwd vs the reference's op chain (cuFFT + einsum + elementwise)."""
    from deno4pytorch_b200.functional import afno_filter
    from oracle import fno_port
    dev = torch.device("cuda", 0)
    gen = torch.Generator().manual_seed(0)
    bs = shape[1] // nb
    x = torch.randn(shape, generator=gen).to(dev).requires_grad_(True)
    p = [torch.view_as_complex(torch.randn(s + (2,), generator=gen) * 0.7 / bs ** 0.5).to(dev).requires_grad_(True)
         for s in ((nb, bs, bs), (nb, bs), (nb, bs, bs), (nb, bs))]
    gy = torch.randn(shape, generator=gen).to(dev)

    def ours():
        afno_filter(x, *p, num_blocks=nb, activation=act).backward(gy)

    def torch_path():
        fno_port.afno_filter(x, *p, num_blocks=nb, activation=act).backward(gy)

    t_ours, t_ref = _time(ours), _time(torch_path)
    A = 4 * x.numel()
    _record(name, {"config": name, "shape": list(shape), "num_blocks": nb, "ours_us": t_ours * 1e3, "torch_gpu_us": t_ref * 1e3,
                   "algo_bytes": 4 * A, "ours_algo_GBps": 4 * A / t_ours / 1e6, "torch_algo_GBps": 4 * A / t_ref / 1e6,
                   "note": "SIMT transforms only; algorithmic bytes = read x, write y, read gy, write gx"})


def test_layer4d_time_vs_torch_fft_path():
    """The 4-D stand-alone layer step (conv_i + w_i + GELU of FNO4d.forward) at the reference script's settings
    (Demo/Turbulence_3d+t/FNO_HIT.py:181-191: modes 16, width 32, batch 2; three time steps) on a 32^3 grid."""
    from deno4pytorch_b200.functional import spectral_conv
    from oracle import fno_port
    dev = torch.device("cuda", 0)
    B, C, spatial, modes = 2, 32, (32, 32, 32, 3), (16, 16, 16, 2)
    gen = torch.Generator().manual_seed(0)
    x = torch.randn((B, C) + spatial, generator=gen).to(dev).requires_grad_(True)
    ws = [(torch.view_as_complex(torch.rand((C, C) + modes + (2,), generator=gen)) / C ** 2).to(dev).requires_grad_(True)
          for _ in range(4)]
    lw = (torch.randn((C, C), generator=gen) / C ** 0.5).to(dev).requires_grad_(True)
    lb = torch.zeros(C, device=dev, requires_grad=True)
    gy = torch.randn((B, C) + spatial, generator=gen).to(dev)

    def ours():
        spectral_conv(x, ws, lw, lb, modes, "gelu").backward(gy)

    def torch_path():
        spec = fno_port.spectral_conv4d(x, ws, modes)
        byp = torch.nn.functional.conv1d(x.reshape(B, C, -1), lw.reshape(C, C, 1), lb).reshape(spec.shape)
        torch.nn.functional.gelu(spec + byp).backward(gy)

    t_ours, t_ref = _time(ours, reps=10, warm=3), _time(torch_path, reps=10, warm=3)
    S = 32 * 32 * 32 * 3
    M = 4 * 16 * 16 * 16 * 2
    algo = 7 * 4 * B * C * S + 3 * 8 * C * C * M
    _record("layer4d", {"config": "layer4d", "shape": [B, C] + list(spatial), "modes": list(modes), "ours_us": t_ours * 1e3,
                        "torch_gpu_us": t_ref * 1e3, "algo_bytes": algo, "ours_algo_GBps": algo / t_ours / 1e6,
                        "torch_algo_GBps": algo / t_ref / 1e6})
