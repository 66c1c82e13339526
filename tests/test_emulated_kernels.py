"""CPU check of the SIMT kernels + C-ABI host logic through the fiber emulation build
(tests/emu/cuda_emu.h).  Same sources nvcc compiles; compares every output and the saved
intermediates with the fp64 truncated-DFT oracle.  The GPU parity tests proper are in
tests/test_gpu_parity.py."""
import numpy as np
import pytest

import emu_harness
from golden_util import golden_names, load_golden, rel_l2
from oracle import dft_truth

TOL = 1e-5  # north_star tolerance for fp32 (relative L2)


def _weights(params):
    return [params[k].numpy() for k in sorted(k for k in params if k.startswith("weights"))]


def _run_case(x, ws, lin_w, lin_b, modes, act, cot):
    layer = emu_harness.EmuLayer(x, ws, lin_w, lin_b, modes, act)
    y = layer.forward(save=True)
    truth = dft_truth.forward(x, ws, lin_w, lin_b, modes, activation=act)
    assert np.isfinite(y).all()
    assert rel_l2(y, truth["y"]) < TOL
    assert rel_l2(layer.z, truth["dact"]) < TOL   # fwd saves act'(z) for the backward pass
    xh = truth["xhat"].reshape(layer.xhat.shape)
    assert rel_l2(layer.xhat, xh) < TOL
    gx, gws, glw, glb = layer.backward(cot)
    tb = dft_truth.backward(x, truth, cot)
    assert np.isfinite(gx).all()
    assert rel_l2(gx, tb["gx"]) < TOL
    for g, t in zip(gws, tb["gweights"]):
        gc = g if np.iscomplexobj(g) else g[..., 0] + 1j * g[..., 1]
        assert rel_l2(gc, t) < TOL
    assert rel_l2(glw.reshape(tb["glin_w"].shape), tb["glin_w"]) < TOL
    assert rel_l2(glb, tb["glin_b"]) < TOL
    return layer


@pytest.mark.parametrize("name", golden_names("layer"))
def test_emulated_layer_matches_truth_on_golden_inputs(name):
    g = load_golden(name)
    m = g["meta"]
    _run_case(g["inputs"]["x"].numpy(), _weights(g["params"]), g["params"]["linear.weight"].numpy(),
              g["params"]["linear.bias"].numpy(), m["modes"], m["activation"] or "silu", g["cot"].numpy())


@pytest.mark.parametrize("name", golden_names("layer"))
def test_emulated_layer_matches_reference_fixture(name):
    """Straight against the reference's own outputs (fp32), not via the oracle."""
    g = load_golden(name)
    m = g["meta"]
    layer = emu_harness.EmuLayer(g["inputs"]["x"].numpy(), _weights(g["params"]), g["params"]["linear.weight"].numpy(),
                                 g["params"]["linear.bias"].numpy(), m["modes"], m["activation"] or "silu")
    y = layer.forward()
    assert rel_l2(y, g["out"]) < TOL
    gx, gws, glw, glb = layer.backward(g["cot"].numpy())
    assert rel_l2(gx, g["grad_inputs"]["x"]) < TOL
    names = sorted(k for k in g["grads"] if k.startswith("weights"))
    for gw, k in zip(gws, names):
        assert rel_l2(gw, g["grads"][k].numpy()) < TOL
    assert rel_l2(glw, g["grads"]["linear.weight"]) < TOL
    assert rel_l2(glb, g["grads"]["linear.bias"]) < TOL


CASES = [
    # spatial, modes, B, cin, cout, act  - shapes chosen to hit tiling edges:
    ((130,), (9,), 3, 5, 3, "gelu"),            # 1-D, two column chunks, ragged channel counts
    ((600,), (16,), 9, 4, 4, "silu"),           # 1-D, several synthesis tiles, ragged plane group
    ((40, 30), (5, 7), 2, 3, 9, "gelu"),        # 2-D, H >= 32: one plane per CTA, cout > 8
    ((37, 21), (12, 11), 1, 2, 2, "tanh"),      # 2-D, odd sizes, max modes on last axis, row tail tile
    ((6, 10), (3, 6), 3, 4, 4, "relu"),         # 2-D tiny plane: several planes per CTA, Nyquist bin
    ((9, 8, 10), (3, 2, 4), 2, 3, 5, "gelu"),   # 3-D
    ((4, 34, 6), (2, 5, 4), 1, 2, 3, "leakyrelu"),  # 3-D, Nyquist on z, H >= 32
    ((16, 12), (2, 3), 2, 1, 1, "identity"),    # single channel
    ((94, 94), (12, 12), 1, 32, 32, "gelu"),    # the Darcy (C2) layer's launch geometry, batch 1
    ((1026,), (16,), 2, 64, 64, "gelu"),        # the Burgers (C1) layer's launch geometry, batch 2
    ((10, 70, 46), (2, 8, 8), 1, 32, 32, "gelu"),  # the C5 plane geometry (70x46 planes), short X
    ((128, 40), (64, 16), 2, 6, 6, "tanh"),     # every row mode retained (2 * modes == H), the Airfoil/Pipe mode counts
    ((34,), (17,), 7, 9, 2, "tanh"),            # 1-D with all N/2+1 modes, cin > cout, odd batch (tiled mix thread tiles ragged)
    ((20, 12), (10, 7), 6, 10, 6, "leakyrelu"),  # 2 * modes == H and all modes on the last axis, 10 -> 6 channels
]


@pytest.mark.parametrize("idx", range(len(CASES)), ids=[str(c[:2]) for c in CASES])
def test_emulated_layer_random_shapes(idx):
    spatial, modes, B, cin, cout, act = CASES[idx]
    rng = np.random.default_rng(100 + idx)
    nd = len(spatial)
    x = rng.standard_normal((B, cin) + spatial).astype(np.float32)
    ws = [((rng.standard_normal((cin, cout) + modes) + 1j * rng.standard_normal((cin, cout) + modes)) / cin)
          .astype(np.complex64) for _ in range(2 ** (nd - 1))]
    lin_w = (rng.standard_normal((cout, cin) + (1,) * nd) / np.sqrt(cin)).astype(np.float32)
    lin_b = rng.standard_normal(cout).astype(np.float32)
    cot = rng.standard_normal((B, cout) + spatial).astype(np.float32)
    _run_case(x, ws, lin_w, lin_b, modes, act, cot)


def test_emulated_isolated_spectral_branch():
    """Bypass zeroed, identity activation: nothing can hide a wrong spectral branch."""
    rng = np.random.default_rng(7)
    spatial, modes = (20, 18), (4, 5)
    x = rng.standard_normal((2, 3) + spatial).astype(np.float32)
    ws = [(rng.standard_normal((3, 4) + modes) + 1j * rng.standard_normal((3, 4) + modes)).astype(np.complex64)
          for _ in range(2)]
    lin_w = np.zeros((4, 3, 1, 1), dtype=np.float32)
    cot = rng.standard_normal((2, 4) + spatial).astype(np.float32)
    _run_case(x, ws, lin_w, None, modes, "identity", cot)


def test_inference_call_skips_saved_tensors():
    rng = np.random.default_rng(3)
    x = rng.standard_normal((2, 3, 12, 10)).astype(np.float32)
    ws = [(rng.standard_normal((3, 3, 2, 3)) + 0j).astype(np.complex64) for _ in range(2)]
    lin_w = rng.standard_normal((3, 3, 1, 1)).astype(np.float32)
    layer = emu_harness.EmuLayer(x, ws, lin_w, None, (2, 3), "gelu")
    y_train = layer.forward(save=True)
    y_inf = layer.forward(save=False)
    assert np.array_equal(y_train, y_inf)
    assert layer.fwd_launches == 3


def test_descriptor_validation_errors():
    from deno4pytorch_b200 import _capi
    rng = np.random.default_rng(0)
    lin_w = np.zeros((2, 2, 1), dtype=np.float32)
    w = [np.zeros((2, 2, 3), dtype=np.complex64)]
    with pytest.raises(_capi.DenoError, match="even length"):
        emu_harness.EmuLayer(rng.standard_normal((1, 2, 15)), w, lin_w, None, 3, "gelu")
    with pytest.raises(_capi.DenoError, match="N/2\\+1"):
        emu_harness.EmuLayer(rng.standard_normal((1, 2, 4)), [np.zeros((2, 2, 4), dtype=np.complex64)], lin_w, None, 4,
                             "gelu")
    lin_w2 = np.zeros((2, 2, 1, 1), dtype=np.float32)
    w2 = [np.zeros((2, 2, 5, 2), dtype=np.complex64)] * 2
    with pytest.raises(_capi.DenoError, match="overlapping"):
        emu_harness.EmuLayer(rng.standard_normal((1, 2, 8, 8)), w2, lin_w2, None, (5, 2), "gelu")


@pytest.mark.parametrize("idx", [2, 5, 8, 12])
def test_emulated_lane_mix_kernel(idx, monkeypatch):
    """mode_mix_lane_kernel (lane = mode, five batch entries per thread, spectrum staged in shared memory) forced for
    every width: same truth, ragged batch / channel tails included."""
    monkeypatch.setenv("DENO_MIX_LANE", "2")
    lib = emu_harness.load()
    lib.deno_reload_env()
    try:
        test_emulated_layer_random_shapes(idx)
    finally:
        monkeypatch.delenv("DENO_MIX_LANE")
        lib.deno_reload_env()


@pytest.mark.parametrize("idx", [0, 2, 5, 8, 11, 13])
def test_emulated_split_role_backward_equals_fused(idx):
    """deno_spectral_bwd_overlap launches the mode mix's input-gradient role and its parameter-gradient (+ DC-mode bias)
    role separately (on the GPU: main stream / side stream).  Same kernels, same summation order: every gradient is
    bit-identical to the single fused launch of deno_spectral_bwd, and one more kernel is launched."""
    spatial, modes, B, cin, cout, act = CASES[idx]
    rng = np.random.default_rng(300 + idx)
    nd = len(spatial)
    x = rng.standard_normal((B, cin) + spatial).astype(np.float32)
    ws = [((rng.standard_normal((cin, cout) + modes) + 1j * rng.standard_normal((cin, cout) + modes)) / cin)
          .astype(np.complex64) for _ in range(2 ** (nd - 1))]
    lin_w = (rng.standard_normal((cout, cin) + (1,) * nd) / np.sqrt(cin)).astype(np.float32)
    lin_b = rng.standard_normal(cout).astype(np.float32)
    cot = rng.standard_normal((B, cout) + spatial).astype(np.float32)
    layer = emu_harness.EmuLayer(x, ws, lin_w, lin_b, modes, act)
    layer.forward(save=True)
    fused = layer.backward(cot)
    n_fused = layer.bwd_launches
    split = layer.backward(cot, split_roles=True)
    assert layer.bwd_launches == n_fused + 1
    assert np.array_equal(fused[0], split[0])
    for a, b in zip(fused[1], split[1]):
        assert np.array_equal(a, b)
    assert np.array_equal(fused[2], split[2]) and np.array_equal(fused[3], split[3])


CASES4 = [
    # spatial (x, y, z, t), modes, B, cin, cout, act
    ((6, 7, 5, 3), (2, 3, 2, 2), 2, 3, 2, "gelu"),        # the fixture's geometry: odd sizes, t = 3 as in the reference
    ((4, 4, 6, 8), (2, 1, 6, 5), 1, 2, 3, "identity"),    # every z mode kept (m3 = Z), Nyquist bin on t, 2 * m1 = X
    ((5, 6, 4, 6), (1, 2, 3, 2), 3, 4, 4, "tanh"),
]


def _truth4(x, ws, lin_w, lin_b, modes, act, cot):
    """float64 torch statement of act(SpectralConv4d(x) + conv1x1(x) + b) and its gradients (oracle/fno_port.py)."""
    import torch
    from oracle import fno_port
    xt = torch.from_numpy(x).double().requires_grad_(True)
    wt = [torch.from_numpy(w).to(torch.complex128).requires_grad_(True) for w in ws]
    lw = torch.from_numpy(lin_w).double().requires_grad_(True)
    lb = torch.from_numpy(lin_b).double().requires_grad_(True)
    B, cin = x.shape[:2]
    spec = fno_port.spectral_conv4d(xt, wt, modes)
    byp = torch.nn.functional.conv1d(xt.reshape(B, cin, -1), lw.reshape(lw.shape[0], cin, 1), lb).reshape(spec.shape)
    y = fno_port.activation_fn(act)(spec + byp)
    (y * torch.from_numpy(cot).double()).sum().backward()
    return y.detach().numpy(), xt.grad.numpy(), [w.grad.numpy() for w in wt], lw.grad.numpy(), lb.grad.numpy()


@pytest.mark.parametrize("idx", range(len(CASES4)))
def test_emulated_4d_layer_matches_float64_reference_algorithm(idx):
    """deno_spectral4_fwd / _bwd (planes (z, t) with a one-sided row axis, two outer-axis passes, four weight blocks)
    against the reference algorithm of Demo/Turbulence_3d+t/FNO_HIT.py:31-78 in float64, forward and every gradient."""
    spatial, modes, B, cin, cout, act = CASES4[idx]
    rng = np.random.default_rng(400 + idx)
    x = rng.standard_normal((B, cin) + spatial).astype(np.float32)
    ws = [((rng.standard_normal((cin, cout) + modes) + 1j * rng.standard_normal((cin, cout) + modes)) / cin)
          .astype(np.complex64) for _ in range(4)]
    lin_w = (rng.standard_normal((cout, cin, 1, 1, 1, 1)) / np.sqrt(cin)).astype(np.float32)
    lin_b = rng.standard_normal(cout).astype(np.float32)
    cot = rng.standard_normal((B, cout) + spatial).astype(np.float32)
    layer = emu_harness.EmuLayer(x, ws, lin_w, lin_b, modes, act)
    y = layer.forward(save=True)
    ty, tgx, tgw, tglw, tglb = _truth4(x, ws, lin_w, lin_b, modes, act, cot)
    assert np.isfinite(y).all()
    assert rel_l2(y, ty) < TOL
    gx, gws, glw, glb = layer.backward(cot)
    assert np.isfinite(gx).all()
    assert rel_l2(gx, tgx) < TOL
    for g, t in zip(gws, tgw):
        assert rel_l2(g, t) < TOL
    assert rel_l2(glw.reshape(tglw.shape), tglw) < TOL
    assert rel_l2(glb, tglb) < TOL


def test_emulated_4d_layer_matches_reference_fixture():
    """The bare SpectralConv4d of the reference (zero bypass, identity) against its own fixture (tests/golden/layer4d.npz)."""
    g = load_golden("layer4d")
    m = g["meta"]
    p = g["params"]
    x = g["inputs"]["x"].numpy()
    ws = [p[f"weights{j}"].numpy() for j in (1, 2, 3, 4)]
    cin, cout = ws[0].shape[:2]
    layer = emu_harness.EmuLayer(x, ws, np.zeros((cout, cin, 1, 1, 1, 1), np.float32), None, tuple(m["modes"]), "identity")
    y = layer.forward()
    assert rel_l2(y, g["out"]) < TOL
    gx, gws, _, _ = layer.backward(g["cot"].numpy())
    assert rel_l2(gx, g["grad_inputs"]["x"]) < TOL
    for j, gw in enumerate(gws):
        assert rel_l2(gw, g["grads"][f"weights{j + 1}"].numpy()) < TOL
