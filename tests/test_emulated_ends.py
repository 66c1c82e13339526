"""CPU checks of the fused lift / projection kernels (deno4pytorch_b200/csrc/fno_pointwise.cuh) through the
host-emulated C ABI: the very kernel sources, compiled with g++ (tests/emu/cuda_emu.h), against the torch-op
formulation of /root/reference/Models/fno/FNOs.py:131-152 evaluated in float64."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from emu_harness import EmuEnds
from golden_util import rel_l2

# (batch, width, hidden, in_feats, out_dim, spatial, padding)
CASES = [
    (2, 8, 16, 1, 1, (19,), 2),            # 1-D, partial last block
    (2, 32, 128, 1, 1, (21, 17), 9),       # the Darcy layout in small
    (1, 16, 24, 3, 2, (13, 15), 0),        # no padding, hidden not a multiple of the reduction round, two outputs
    (1, 20, 32, 10, 1, (5, 6, 7), 3),      # 3-D roll-out style input (10 steps)
]


def _torch_reference(case, seed):
    B, Cw, Hd, Kx, O, spatial, pad = case
    nd = len(spatial)
    g = torch.Generator().manual_seed(seed)
    dd = torch.float64
    x = torch.randn((B,) + spatial + (Kx,), generator=g, dtype=dd).requires_grad_(True)
    grid = torch.rand((B,) + spatial + (nd,), generator=g, dtype=dd)
    w0 = (torch.randn(Cw, Kx + nd, generator=g, dtype=dd) * 0.5).requires_grad_(True)
    b0 = (torch.randn(Cw, generator=g, dtype=dd) * 0.1).requires_grad_(True)
    w1 = (torch.randn(Hd, Cw, generator=g, dtype=dd) / Cw ** 0.5).requires_grad_(True)
    b1 = (torch.randn(Hd, generator=g, dtype=dd) * 0.1).requires_grad_(True)
    w2 = (torch.randn(O, Hd, generator=g, dtype=dd) / Hd ** 0.5).requires_grad_(True)
    b2 = (torch.randn(O, generator=g, dtype=dd) * 0.1).requires_grad_(True)
    return dict(x=x, grid=grid, w0=w0, b0=b0, w1=w1, b1=b1, w2=w2, b2=b2), g


def _np(t):
    return t.detach().numpy().astype(np.float32)


@pytest.mark.parametrize("case", CASES, ids=[str(c[5]) + f"-w{c[1]}" for c in CASES])
def test_lift_forward_and_backward_match_torch_ops(case):
    B, Cw, Hd, Kx, O, spatial, pad = case
    nd = len(spatial)
    t, g = _torch_reference(case, 1)
    h = F.linear(torch.cat((t["x"], t["grid"]), -1), t["w0"], t["b0"]).movedim(-1, 1)
    if pad:
        h = F.pad(h, [0, pad] * nd)
    gh = torch.randn(h.shape, generator=g, dtype=torch.float64)
    h.backward(gh)
    e = EmuEnds(B, Cw, Hd, Kx, nd, O, spatial, pad)
    assert e.supported()
    got = e.lift(_np(t["x"]), _np(t["grid"]), _np(t["w0"]), _np(t["b0"]))
    assert got.shape == tuple(h.shape)
    assert rel_l2(got, h.detach().numpy()) < 2e-6
    if pad:
        assert np.all(got[(Ellipsis,) + (slice(-pad, None),)] == 0.0)      # the padding is written, as zeros
    gx, gw0, gb0 = e.lift_bwd(_np(gh), _np(t["x"]), _np(t["grid"]), _np(t["w0"]))
    assert rel_l2(gx, t["x"].grad.numpy()) < 2e-6
    assert rel_l2(gw0, t["w0"].grad.numpy()) < 2e-6
    assert rel_l2(gb0, t["b0"].grad.numpy()) < 2e-6
    gx2, gw0b, _ = e.lift_bwd(_np(gh), _np(t["x"]), _np(t["grid"]), _np(t["w0"]), want_gx=False)
    assert gx2 is None and np.array_equal(gw0b, gw0)


@pytest.mark.parametrize("case", CASES, ids=[str(c[5]) + f"-w{c[1]}" for c in CASES])
def test_projection_forward_and_backward_match_torch_ops(case):
    B, Cw, Hd, Kx, O, spatial, pad = case
    nd = len(spatial)
    t, g = _torch_reference(case, 2)
    padded = tuple(s + pad for s in spatial)
    hp = torch.randn((B, Cw) + padded, generator=g, dtype=torch.float64).requires_grad_(True)
    h = hp[(Ellipsis,) + (slice(None, -pad),) * nd] if pad else hp
    y = F.linear(F.gelu(F.linear(h.movedim(1, -1), t["w1"], t["b1"])), t["w2"], t["b2"])
    gy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    y.backward(gy)
    e = EmuEnds(B, Cw, Hd, Kx, nd, O, spatial, pad)
    got = e.proj(_np(hp), _np(t["w1"]), _np(t["b1"]), _np(t["w2"]), _np(t["b2"]))
    assert got.shape == tuple(y.shape)
    assert rel_l2(got, y.detach().numpy()) < 2e-6
    gh, gw1, gb1, gw2, gb2 = e.proj_bwd(_np(gy), _np(hp), _np(t["w1"]), _np(t["b1"]), _np(t["w2"]))
    assert rel_l2(gh, hp.grad.numpy()) < 3e-6
    if pad:
        assert np.all(gh[(Ellipsis,) + (slice(-pad, None),)] == 0.0)
    assert rel_l2(gw1, t["w1"].grad.numpy()) < 3e-6
    assert rel_l2(gb1, t["b1"].grad.numpy()) < 3e-6
    assert rel_l2(gw2, t["w2"].grad.numpy()) < 3e-6
    assert rel_l2(gb2, t["b2"].grad.numpy()) < 3e-6


def test_unsupported_descriptors_are_reported():
    assert not EmuEnds(1, 30, 16, 1, 2, 1, (8, 8), 0).supported()      # width not built
    assert not EmuEnds(1, 32, 16, 15, 2, 1, (8, 8), 0).supported()     # 17 input features
    assert not EmuEnds(1, 32, 16, 1, 2, 5, (8, 8), 0).supported()      # five outputs
