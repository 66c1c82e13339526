"""The selectable kernels behind the measured-best defaults stay correct: every switch that picks ANOTHER kernel for a
stage (SIMT row synthesis, one-buffer analysis staging, ring without staged act'(z), tensor-core / lane mode mix,
register-staged weight gradient, one-stream backward) against the reference algorithm in float64 on the GPU
(oracle/fno_port.py: /root/reference/Models/fno/spectral_layers.py:183-208, 299-338), forward and every gradient.
Tolerance 1e-5 relative L2 (north_star)."""
import pytest
import torch

from golden_util import rel_l2

pytestmark = pytest.mark.gpu

TOL = 1e-5

SWITCHES = [
    ("DENO_ROWSYNTH_TC", "0"),     # SIMT row_synthesis instead of the tcgen05 GEMM
    ("DENO_AN_RING", "0"),         # one staging buffer, act'(z) by per-thread loads
    ("DENO_AN_RING", "2"),         # ring, act'(z) by per-thread loads (separate gz pass)
    ("DENO_MIX_LANE", "0"),        # tensor-core mix also at width 128
    ("DENO_MIX_LANE", "2"),        # lane mix at every width
    ("DENO_MIX_TC", "0"),          # SIMT mix at width >= 64
    ("DENO_WGRAD_TMA", "0"),       # register-staged bypass weight gradient
]
SHAPES = [
    # B, C, spatial, modes
    (5, 32, (94, 94), (12, 12)),        # Darcy plane (tight shared-memory fit of the ring), odd batch tail of the lane kernel
    (3, 64, (72, 72), (12, 12)),        # width 64: tcgen05 mix by default
    (2, 128, (41, 36), (6, 5)),         # width 128: lane mix by default, odd height, odd mode count on the last axis
    (2, 32, (10, 70, 46), (3, 8, 8)),   # 3-D planes of the turbulence config, short outer axis
]


@pytest.mark.parametrize("shape", SHAPES, ids=[f"w{s[1]}-{'x'.join(map(str, s[2]))}" for s in SHAPES])
@pytest.mark.parametrize("switch", SWITCHES, ids=[f"{k}={v}" for k, v in SWITCHES])
def test_alternative_kernels_match_float64_reference(switch, shape, monkeypatch):
    from conftest import set_deno_env
    from deno4pytorch_b200.functional import spectral_conv
    from oracle import fno_port
    set_deno_env(monkeypatch, switch[0], switch[1])
    B, C, spatial, modes = shape
    nd = len(spatial)
    dev = torch.device("cuda", 0)
    gen = torch.Generator().manual_seed(41)
    x = torch.randn((B, C) + spatial, generator=gen)
    ws = [torch.view_as_complex(torch.randn((C, C) + modes + (2,), generator=gen) / C) for _ in range(2 ** (nd - 1))]
    lw = torch.randn((C, C) + (1,) * nd, generator=gen) / C ** 0.5
    lb = torch.randn(C, generator=gen)
    cot = torch.randn((B, C) + spatial, generator=gen)

    xg = x.to(dev).requires_grad_(True)
    wg = [w.to(dev).requires_grad_(True) for w in ws]
    lwg, lbg = lw.to(dev).requires_grad_(True), lb.to(dev).requires_grad_(True)
    y = spectral_conv(xg, wg, lwg, lbg, modes, "gelu")
    y.backward(cot.to(dev))
    torch.cuda.synchronize()

    xr = x.to(dev).double().requires_grad_(True)
    wr = [w.to(dev).to(torch.complex128).requires_grad_(True) for w in ws]
    lwr, lbr = lw.to(dev).double().requires_grad_(True), lb.to(dev).double().requires_grad_(True)
    ref = fno_port.spectral_conv(xr, wr, lwr, lbr, modes, activation="gelu")
    ref.backward(cot.to(dev).double())
    torch.cuda.synchronize()

    assert rel_l2(y.detach().cpu(), ref.detach().cpu()) < TOL
    assert rel_l2(xg.grad.cpu(), xr.grad.cpu()) < TOL
    for j, (a, b) in enumerate(zip(wg, wr)):
        assert rel_l2(a.grad.cpu(), b.grad.cpu()) < TOL, f"weights{j + 1}"
    assert rel_l2(lwg.grad.cpu(), lwr.grad.cpu()) < TOL
    assert rel_l2(lbg.grad.cpu(), lbr.grad.cpu()) < TOL


def test_one_stream_backward_switch_matches_default():
    """DENO_BWD_OVERLAP=0 / functional.set_bwd_overlap(False): same kernels on one stream, bit-identical gradients
    (the per-shape bitwise test lives in test_gpu_parity.py; this one goes through a whole model)."""
    import deno4pytorch_b200 as d4
    from deno4pytorch_b200 import functional
    dev = torch.device("cuda", 0)
    torch.manual_seed(5)
    model = d4.FNO2d(in_dim=1, out_dim=1, modes=(6, 6), width=32, depth=2, padding=3).to(dev)
    x = torch.randn(3, 29, 31, 1, device=dev)
    grid = torch.rand(3, 29, 31, 2, device=dev)
    tgt = torch.randn(3, 29, 31, 1, device=dev)
    grads = {}
    try:
        for on in (True, False):
            functional.set_bwd_overlap(on)
            model.zero_grad(set_to_none=True)
            torch.nn.functional.mse_loss(model(x, grid), tgt).backward()
            torch.cuda.synchronize()
            grads[on] = [(torch.view_as_real(p.grad) if p.grad.is_complex() else p.grad).clone() for p in model.parameters()]
    finally:
        functional.set_bwd_overlap(True)
    for a, b in zip(grads[True], grads[False]):
        assert torch.equal(a, b)
