"""The layers' SECOND caller: ``SpectralRegressor``
(/root/reference/Models/transformer/Transformers.py:346-470).  That module imports the layers with
``from fno.spectral_layers import *`` (Transformers.py:25), so with this repository's ``fno`` package first on the path it
runs on the CUDA kernels unchanged.  Its use of the layers differs from the FNO stacks in every way the layer API allows:
real-view (``use_complex=False``) weights - also in 2-D, the variant the reference's own layer crashes on -, the default
``norm='ortho'``, ``silu``, ``in_dim != out_dim``, the last layer's activation swapped for an ``Identity`` module, dropout.
The test restates the regressor's forward (permute, layers, permute, two-layer MLP; lines 417-470) over this repo's layers
on the GPU and over the CPU oracle with the same parameters, and compares the output, the input gradient and EVERY
parameter gradient.  Tolerance 1e-5 relative L2."""
import pytest
import torch
import torch.nn as nn

from golden_util import rel_l2

pytestmark = pytest.mark.gpu

TOL = 1e-5


class _Identity(nn.Module):          # the reference's basic.basic_layers.Identity: forward(x) -> x
    def forward(self, x):
        return x


class _Regressor(nn.Module):
    """``SpectralRegressor(in_dim, n_hidden, freq_dim, out_dim, modes, num_spectral_layers, spacial_dim, activation,
    last_activation, dropout)`` with ``spacial_fc=False, normalizer=None`` (the transformer demos' setting)."""

    def __init__(self, layer_cls, n_hidden, freq_dim, out_dim, modes, nlayers, spacial_dim, activation, last_activation, dropout):
        super().__init__()
        from deno4pytorch_b200.configs import activation_dict
        self.spectral_conv = nn.ModuleList(
            [layer_cls(in_dim=n_hidden, out_dim=freq_dim, modes=modes, dropout=dropout, activation=activation, use_complex=False)])
        for _ in range(nlayers - 1):
            self.spectral_conv.append(
                layer_cls(in_dim=freq_dim, out_dim=freq_dim, modes=modes, dropout=dropout, activation=activation, use_complex=False))
        if not last_activation:
            self.spectral_conv[-1].activation = _Identity()
        dff = spacial_dim * freq_dim
        self.regressor = nn.Sequential(nn.Linear(freq_dim, dff), activation_dict[activation], nn.Linear(dff, out_dim))

    def forward(self, x):
        nd = x.dim() - 2
        x = x.movedim(-1, 1)
        for layer in self.spectral_conv:
            x = layer(x)
        x = x.movedim(1, -1)
        assert x.dim() == nd + 2
        return self.regressor(x)


def _leaf64(t):
    return t.detach().cpu().double().requires_grad_(True)


def _oracle_params(model):
    """float64 leaf copies of every parameter the regressor owns, keyed like ``model.named_parameters()``."""
    return {k: _leaf64(p) for k, p in model.named_parameters()}


def _oracle_forward(model, p64, x, modes, activation, last_activation):
    from oracle import fno_port
    h = x.movedim(-1, 1)
    n = len(model.spectral_conv)
    for i, layer in enumerate(model.spectral_conv):
        nblk = len(layer._weights())
        ws = [torch.view_as_complex(p64[f"spectral_conv.{i}.weights{j + 1}"]) for j in range(nblk)]   # real-view leaves
        act = activation if (last_activation or i + 1 < n) else "identity"
        h = fno_port.spectral_conv(h, ws, p64[f"spectral_conv.{i}.linear.weight"], p64[f"spectral_conv.{i}.linear.bias"],
                                   modes, activation=act, norm="ortho")
    h = h.movedim(1, -1)
    _, actm, _ = model.regressor
    h = torch.nn.functional.linear(h, p64["regressor.0.weight"], p64["regressor.0.bias"])
    h = actm(h)
    return torch.nn.functional.linear(h, p64["regressor.2.weight"], p64["regressor.2.bias"])


CASES = [
    # spacial_dim, spatial, n_hidden, freq_dim, out_dim, modes, layers, activation, last_activation
    (1, (64,), 24, 16, 1, 12, 2, "silu", True),
    (2, (33, 30), 20, 12, 2, (8, 9), 2, "silu", False),     # 2-D real-view weights + Identity swap
    (3, (12, 10, 14), 8, 8, 1, (3, 4, 5), 3, "gelu", True),
]


@pytest.mark.parametrize("case", CASES, ids=["1d", "2d", "3d"])
def test_regressor_composition_matches_oracle(case):
    import deno4pytorch_b200 as d4
    sd, spatial, n_hidden, freq_dim, out_dim, modes, nlayers, activation, last_act = case
    layer_cls = {1: d4.SpectralConv1d, 2: d4.SpectralConv2d, 3: d4.SpectralConv3d}[sd]
    dev = torch.device("cuda", 0)
    torch.manual_seed(7)
    model = _Regressor(layer_cls, n_hidden, freq_dim, out_dim, modes, nlayers, sd, activation, last_act, dropout=0.1).to(dev)
    model.eval()                                         # dropout off: the deterministic path
    for layer in model.spectral_conv:                    # the reference's init scale (1 / (in * out)) makes the spectral
        for w in layer._weights():                       # branch negligible: scale it up so that the check means something
            w.data.mul_(float(layer.in_dim))
        assert not any(w.is_complex() for w in layer._weights())      # real-view parameters, as SpectralRegressor asks
    g = torch.Generator().manual_seed(8)
    x = torch.randn((3,) + spatial + (n_hidden,), generator=g)
    cot = torch.randn((3,) + spatial + (out_dim,), generator=g)
    xg = x.to(dev).requires_grad_(True)
    y = model(xg)
    (y * cot.to(dev)).sum().backward()
    torch.cuda.synchronize()
    x64 = x.double().requires_grad_(True)
    p64 = _oracle_params(model)
    ref = _oracle_forward(model, p64, x64, modes, activation, last_act)
    (ref * cot.double()).sum().backward()
    assert y.shape == ref.shape
    assert rel_l2(y.detach().cpu(), ref.detach()) < TOL
    assert rel_l2(xg.grad.cpu(), x64.grad) < TOL
    # every parameter gradient: the real-view spectral weights, the bypass convs, the regressor MLP
    for k, p in model.named_parameters():
        assert p.grad is not None, k
        assert rel_l2(p.grad.cpu(), p64[k].grad) < TOL, k
    # training mode with dropout p > 0 runs (the 2-D layer drops the spectral branch's input, spectral_layers.py:189-191)
    model.train()
    out = model(xg.detach())
    assert torch.isfinite(out).all()
