"""GPU parity of the 4-D stand-alone layer / model (SpectralConv4d, FNO4d: /root/reference/Demo/Turbulence_3d+t/FNO_HIT.py:31-171)
through deno_spectral4_fwd / _bwd: the reference's own fixtures (tests/golden/layer4d.npz, fno4d.npz, written by
tests/golden/make_golden_4d.py from the reference's class definitions) and the float64 reference algorithm on the GPU at a
tensor-core-eligible width, both kernel families.  Tolerance 1e-5 relative L2 (north_star)."""
import pytest
import torch

from golden_util import load_golden, rel_l2

pytestmark = pytest.mark.gpu

TOL = 1e-5


@pytest.fixture(params=["simt", "auto"], autouse=True)
def spectral_path(request, monkeypatch):
    from conftest import set_deno_env
    set_deno_env(monkeypatch, "DENO_SPECTRAL_PATH", request.param)
    return request.param


def _dev():
    return torch.device("cuda", 0)


def test_layer4d_matches_reference_fixture():
    from deno4pytorch_b200.FNO_HIT import SpectralConv4d
    g = load_golden("layer4d")
    m = g["meta"]
    layer = SpectralConv4d(m["in_dim"], m["out_dim"], *m["modes"]).to(_dev())
    layer.load_state_dict({k: v for k, v in g["params"].items()}, strict=True)
    x = g["inputs"]["x"].to(_dev()).requires_grad_(True)
    y = layer(x)
    (y * g["cot"].to(_dev())).sum().backward()
    assert rel_l2(y.detach().cpu(), g["out"]) < TOL
    assert rel_l2(x.grad.cpu(), g["grad_inputs"]["x"]) < TOL
    for k, p in layer.named_parameters():
        assert rel_l2(p.grad.cpu(), g["grads"][k]) < TOL, k


def test_fno4d_matches_reference_fixture_and_state_dict_roundtrip():
    from deno4pytorch_b200.FNO_HIT import FNO4d
    g = load_golden("fno4d")
    m = g["meta"]
    model = FNO4d(*m["modes"][:3], 4, m["width"]).to(_dev())       # modes4 = 4 is clamped to 2 like the reference does
    assert list(model.state_dict().keys()) == list(g["params"].keys())
    model.load_state_dict(g["params"], strict=True)
    x = g["inputs"]["x"].to(_dev()).requires_grad_(True)
    y = model(x)
    (y * g["cot"].to(_dev())).sum().backward()
    assert rel_l2(y.detach().cpu(), g["out"]) < TOL
    assert rel_l2(x.grad.cpu(), g["grad_inputs"]["x"]) < TOL
    for k, p in model.named_parameters():
        assert rel_l2(p.grad.cpu(), g["grads"][k]) < 2 * TOL, k     # fc0's gradient sums 2 k tiny terms of both signs


@pytest.mark.parametrize("shape", [
    (2, 32, 32, (6, 8, 16, 6), (3, 2, 5, 3)),     # width 32: the tensor-core plane kernels on the default path
    (1, 8, 16, (4, 5, 7, 3), (2, 2, 7, 2)),       # in != out, every z mode kept, the reference's 3-step time axis
], ids=["w32", "in8out16"])
def test_layer4d_matches_float64_reference_algorithm(shape):
    from deno4pytorch_b200.functional import spectral_conv
    from oracle import fno_port
    B, cin, cout, spatial, modes = shape
    dev = _dev()
    gen = torch.Generator().manual_seed(51)
    x = torch.randn((B, cin) + spatial, generator=gen)
    ws = [torch.view_as_complex(torch.randn((cin, cout) + modes + (2,), generator=gen) / cin) for _ in range(4)]
    lw = torch.randn((cout, cin), generator=gen) / cin ** 0.5
    lb = torch.randn(cout, generator=gen)
    cot = torch.randn((B, cout) + spatial, generator=gen)

    xg = x.to(dev).requires_grad_(True)
    wg = [w.to(dev).requires_grad_(True) for w in ws]
    lwg, lbg = lw.to(dev).requires_grad_(True), lb.to(dev).requires_grad_(True)
    y = spectral_conv(xg, wg, lwg, lbg, modes, "gelu")
    y.backward(cot.to(dev))
    torch.cuda.synchronize()

    xr = x.to(dev).double().requires_grad_(True)
    wr = [w.to(dev).to(torch.complex128).requires_grad_(True) for w in ws]
    lwr, lbr = lw.to(dev).double().requires_grad_(True), lb.to(dev).double().requires_grad_(True)
    spec = fno_port.spectral_conv4d(xr, wr, modes)
    byp = torch.nn.functional.conv1d(xr.reshape(B, cin, -1), lwr.reshape(cout, cin, 1), lbr).reshape(spec.shape)
    ref = torch.nn.functional.gelu(spec + byp)
    ref.backward(cot.to(dev).double())
    torch.cuda.synchronize()

    assert rel_l2(y.detach().cpu(), ref.detach().cpu()) < TOL
    assert rel_l2(xg.grad.cpu(), xr.grad.cpu()) < TOL
    for j, (a, b) in enumerate(zip(wg, wr)):
        assert rel_l2(a.grad.cpu(), b.grad.cpu()) < TOL, f"weights{j + 1}"
    assert rel_l2(lwg.grad.cpu(), lwr.grad.cpu()) < TOL
    assert rel_l2(lbg.grad.cpu(), lbr.grad.cpu()) < TOL
