"""Pins both CPU oracles against fixtures generated from the real reference
(tests/golden/make_golden.py) - CPU only."""
import os
import sys

import numpy as np
import pytest
import torch

from golden_util import golden_names, load_golden, rel_l2
from oracle import dft_truth, fno_port

LAYERS = golden_names("layer")
MODELS = golden_names("model")
TOL = 2e-6  # fp32 reference output vs fp32 port / fp64 truth; reference noise floor is ~1e-7


def _layer_weights(params):
    return [params[k] for k in sorted(k for k in params if k.startswith("weights"))]


def test_fixture_inventory():
    assert len(LAYERS) >= 9 and len(MODELS) >= 5
    assert {load_golden(n)["meta"]["ndim"] for n in LAYERS} == {1, 2, 3}


@pytest.mark.parametrize("name", LAYERS)
def test_port_layer_matches_reference(name):
    g = load_golden(name)
    m = g["meta"]
    x = g["inputs"]["x"].clone().requires_grad_(True)
    params = {k: v.clone().requires_grad_(True) for k, v in g["params"].items()}
    y = fno_port.spectral_conv(x, _layer_weights(params), params["linear.weight"], params["linear.bias"],
                               m["modes"], activation=m["activation"], norm=m["norm"])
    assert rel_l2(y, g["out"]) < TOL
    (y * g["cot"]).sum().backward()
    assert rel_l2(x.grad, g["grad_inputs"]["x"]) < TOL
    for k, p in params.items():
        assert rel_l2(p.grad, g["grads"][k]) < TOL, k


@pytest.mark.parametrize("name", MODELS)
def test_port_model_matches_reference(name):
    g = load_golden(name)
    m = g["meta"]
    x = g["inputs"]["x"].clone().requires_grad_(True)
    params = {k: v.clone().requires_grad_(True) for k, v in g["params"].items()}
    y = fno_port.fno_forward(params, x, g["inputs"]["grid"], ndim=m["ndim"], modes=m["modes"], depth=m["depth"],
                             padding=m["padding"], activation=m.get("activation", "gelu"))
    assert y.shape == g["out"].shape
    assert rel_l2(y, g["out"]) < TOL
    (y * g["cot"]).sum().backward()
    assert rel_l2(x.grad, g["grad_inputs"]["x"]) < 5e-6
    for k, p in params.items():
        assert rel_l2(p.grad, g["grads"][k]) < 5e-6, k


@pytest.mark.parametrize("name", LAYERS)
def test_truth_layer_matches_reference(name):
    """The truncated-DFT math contract (what the CUDA kernels implement) reproduces the
    reference's FFT path, forward and all five gradients."""
    g = load_golden(name)
    m = g["meta"]
    x = g["inputs"]["x"].numpy()
    ws = [w.numpy() for w in _layer_weights(g["params"])]
    fwd = dft_truth.forward(x, ws, g["params"]["linear.weight"].numpy(), g["params"]["linear.bias"].numpy(),
                            m["modes"], activation=m["activation"])
    assert rel_l2(fwd["y"], g["out"]) < TOL
    bwd = dft_truth.backward(x, fwd, g["cot"].numpy())
    assert rel_l2(bwd["gx"], g["grad_inputs"]["x"]) < TOL
    names = sorted(k for k in g["grads"] if k.startswith("weights"))
    for k, gw in zip(names, bwd["gweights"]):
        ref = g["grads"][k]
        if not ref.is_complex():
            ref = torch.view_as_complex(ref.contiguous())
        assert rel_l2(torch.from_numpy(gw), ref) < TOL, k
    assert rel_l2(bwd["glin_w"].reshape(g["grads"]["linear.weight"].shape), g["grads"]["linear.weight"]) < TOL
    assert rel_l2(bwd["glin_b"], g["grads"]["linear.bias"]) < TOL


def test_truth_vs_port_fp64_isolated_spectral_branch():
    """Isolated spectral branch (bypass zeroed, identity activation), O(1) weights - the check
    SURVEY section 4 asks for so that a wrong spectral branch cannot hide behind the bypass."""
    rng = np.random.default_rng(0)
    for spatial, modes in [((16,), (5,)), ((9, 12), (3, 4)), ((6, 7, 8), (2, 3, 4))]:
        nd = len(spatial)
        cin, cout, b = 3, 4, 2
        x = rng.standard_normal((b, cin) + spatial)
        ws = [rng.standard_normal((cin, cout) + modes) + 1j * rng.standard_normal((cin, cout) + modes)
              for _ in range(2 ** (nd - 1))]
        lw = np.zeros((cout, cin) + (1,) * nd)
        fwd = dft_truth.forward(x, ws, lw, None, modes, activation="identity")
        xt = torch.from_numpy(x).requires_grad_(True)
        wt = [torch.from_numpy(w).requires_grad_(True) for w in ws]
        y = fno_port.spectral_conv(xt, wt, torch.from_numpy(lw), None, modes, activation="identity")
        assert rel_l2(fwd["y"], y) < 1e-13
        cot = rng.standard_normal(y.shape)
        (y * torch.from_numpy(cot)).sum().backward()
        bwd = dft_truth.backward(x, fwd, cot)
        assert rel_l2(bwd["gx"], xt.grad) < 1e-13
        for gw, w in zip(bwd["gweights"], wt):
            assert rel_l2(torch.from_numpy(gw), w.grad) < 1e-13


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree only exists in the build container")
def test_port_matches_live_reference_darcy_shape():
    """Direct check against the imported reference at the C2 (Darcy) layer shape, reduced batch."""
    from golden_util import reference_modules
    with reference_modules() as (_, ref_fnos):
        torch.manual_seed(0)
        ref = ref_fnos.FNO2d(in_dim=1, out_dim=1, modes=(12, 12), width=32, depth=4, steps=1, padding=9,
                             activation="gelu")
        x = torch.randn(2, 85, 85, 1)
        grid = fno_port.make_grid(2, (85, 85))
        want = ref(x, grid)
        params = {k: v.detach() for k, v in ref.state_dict().items()}
    got = fno_port.fno_forward(params, x, grid, ndim=2, modes=(12, 12), depth=4, padding=9)
    assert rel_l2(got, want) < TOL


# ---- 4-D stand-alone layer / model (FNO_HIT.py): the oracle is pinned first ----------------
def test_port_layer4d_matches_reference():
    g = load_golden("layer4d")
    m = g["meta"]
    assert m["kind"] == "layer4d" and m["modes"][3] == 2     # the reference clamps modes4 to 3 // 2 + 1
    x = g["inputs"]["x"].clone().requires_grad_(True)
    params = {k: v.clone().requires_grad_(True) for k, v in g["params"].items()}
    y = fno_port.spectral_conv4d(x, [params[f"weights{j}"] for j in (1, 2, 3, 4)], m["modes"])
    assert rel_l2(y, g["out"]) < TOL
    (y * g["cot"]).sum().backward()
    assert rel_l2(x.grad, g["grad_inputs"]["x"]) < TOL
    for k, p in params.items():
        assert rel_l2(p.grad, g["grads"][k]) < TOL, k


def test_port_layer4d_keeps_only_the_low_block_on_the_third_axis():
    """Unlike the 3-D layer, SpectralConv4d drops the negative frequencies of its third axis (FNO_HIT.py:63-72 slice
    ``:modes3`` only): a pure negative-frequency wave along z must come out as zero, a positive one must not."""
    torch.manual_seed(0)
    Z = 8
    z = torch.arange(Z, dtype=torch.float32)
    w = [torch.ones(1, 1, 1, 1, 3, 1, dtype=torch.complex64) for _ in range(4)]
    # real input cannot isolate one sign of a NON-last axis unless another axis carries the conjugate: use x-axis phase
    X = 6
    xs = torch.arange(X, dtype=torch.float32)
    pos = torch.cos(2 * torch.pi * (xs[:, None] / X * 0 + 2 * z[None, :] / Z))          # k_z = +-2 (both signs, k_x = 0)
    field = pos[None, None, :, None, :, None].expand(1, 1, X, 4, Z, 3).contiguous()
    y = fno_port.spectral_conv4d(field, w, (1, 1, 3, 1))
    # only the +2 half survives: amplitude halves and the result is complex-analytic in z -> real part cos/2 after C2R
    assert torch.allclose(y[0, 0, :, 0, :, 0], 0.5 * pos, atol=1e-5)


def test_port_fno4d_matches_reference():
    g = load_golden("fno4d")
    m = g["meta"]
    assert m["kind"] == "model4d"
    x = g["inputs"]["x"].clone().requires_grad_(True)
    params = {k: v.clone().requires_grad_(True) for k, v in g["params"].items()}
    y = fno_port.fno4d_forward(params, x, m["modes"])
    assert rel_l2(y, g["out"]) < TOL
    (y * g["cot"]).sum().backward()
    assert rel_l2(x.grad, g["grad_inputs"]["x"]) < TOL
    for k, p in params.items():
        assert rel_l2(p.grad, g["grads"][k]) < 5 * TOL, k


# ---- adaptive Fourier layers (spectral_layers.py:341-573) ------------------------------
AFNO = golden_names("afno")


def test_afno_fixture_inventory():
    assert {load_golden(n)["meta"]["ndim"] for n in AFNO} == {1, 2, 3} and len(AFNO) >= 5


@pytest.mark.parametrize("name", AFNO)
def test_port_afno_matches_reference(name):
    g = load_golden(name)
    m = g["meta"]
    x = g["inputs"]["x"].clone().requires_grad_(True)
    p = {k: v.clone().requires_grad_(True) for k, v in g["params"].items()}
    y = fno_port.afno_filter(x, p["weights1"], p["bias1"], p["weights2"], p["bias2"], num_blocks=m["num_blocks"],
                             sparsity_threshold=m["sparsity_threshold"],
                             hard_thresholding_fraction=m["hard_thresholding_fraction"], activation=m["activation"])
    assert rel_l2(y, g["out"]) < TOL
    (y * g["cot"]).sum().backward()
    assert rel_l2(x.grad, g["grad_inputs"]["x"]) < TOL
    for k, v in p.items():
        assert rel_l2(v.grad, g["grads"][k]) < TOL, k


def test_afno_kept_modes_follows_the_reference_slice():
    # 2-D / 3-D: total_modes counts the whole grid, so fraction 1 keeps the full half-spectrum axis
    assert fno_port.afno_kept_modes((55, 64), 1) == 33
    assert fno_port.afno_kept_modes((6, 9), 0.1) == 2          # int(28 * 0.1)
    assert fno_port.afno_kept_modes((16,), 0.5) == 4            # int(9 * 0.5)
    assert fno_port.afno_kept_modes((15,), 1) == 8
    assert fno_port.afno_kept_modes((4, 4), 0.01) == 0
