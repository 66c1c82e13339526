"""CPU check of the adaptive Fourier layer path (include/deno_afno.h: full-spectrum SIMT transforms + the fused block
MLP kernels of csrc/afno_pointwise.cuh) through the fiber emulation build: the same sources nvcc compiles, against the
reference's own fixtures (tests/golden/afno*.npz) and the float64 reference algorithm."""
import numpy as np
import pytest
import torch

import emu_harness
from golden_util import golden_names, load_golden, rel_l2
from oracle import fno_port

TOL = 1e-5  # north_star tolerance for fp32 (relative L2)


def _layer(g):
    m, p = g["meta"], g["params"]
    x = g["inputs"]["x"].numpy()
    kept = fno_port.afno_kept_modes(x.shape[2:], m["hard_thresholding_fraction"])
    return emu_harness.EmuAfno(x, p["weights1"].numpy(), p["bias1"].numpy(), p["weights2"].numpy(), p["bias2"].numpy(),
                               m["num_blocks"], kept, m["activation"], m["sparsity_threshold"])


@pytest.mark.parametrize("name", golden_names("afno"))
def test_emulated_afno_matches_reference_fixture(name):
    g = load_golden(name)
    layer = _layer(g)
    y = layer.forward()
    assert np.isfinite(y).all() and np.isfinite(layer.xhat.view(np.float32)).all()
    assert rel_l2(y, g["out"]) < TOL
    gx, gp = layer.backward(g["cot"].numpy())
    assert np.isfinite(gx).all()
    assert rel_l2(gx, g["grad_inputs"]["x"]) < TOL
    for got, k in zip(gp, ("weights1", "bias1", "weights2", "bias2")):
        assert rel_l2(got, g["grads"][k].numpy()) < TOL, k


def test_emulated_afno_saved_spectrum_is_the_ortho_rfftn():
    g = load_golden("afno2d")
    layer = _layer(g)
    layer.forward()
    x = g["inputs"]["x"]
    want = torch.fft.rfft2(x.double(), norm="ortho").numpy()          # (B, C, H, W/2+1): every row frequency, f_k = k
    assert rel_l2(layer.xhat.reshape(want.shape), want) < TOL


def _reference_f64(x, p, num_blocks, lam, frac, act, cot):
    xr = torch.from_numpy(x).double().requires_grad_(True)
    pr = [torch.from_numpy(a).to(torch.complex128).requires_grad_(True) for a in p]
    y = fno_port.afno_filter(xr, pr[0], pr[1], pr[2], pr[3], num_blocks=num_blocks, sparsity_threshold=lam,
                             hard_thresholding_fraction=frac, activation=act)
    y.backward(torch.from_numpy(cot).double())
    return y.detach().numpy(), xr.grad.numpy(), [q.grad.numpy() for q in pr]


@pytest.mark.parametrize("case", [
    # shape, blocks, lambda, fraction, act:  several tiles per batch entry + a ragged last tile, block size not a
    # multiple of 4, one block of 12 (pairs spread over the threads), 3-D with odd sizes
    ((2, 6, 9, 40), 2, 0.02, 1, "gelu"),
    ((1, 12, 70), 1, 0.01, 1, "relu"),
    ((2, 10, 24), 5, 0.05, 0.6, "tanh"),
    ((1, 4, 3, 5, 12), 2, 0.01, 1, "leakyrelu"),
    ((3, 8, 5, 6), 8, 0.0, 1, "identity"),
    # the largest blocks: 64 (every pair slot of the backward's 17 per thread in use, 64-bin tiles) and 32
    ((2, 64, 6, 40), 1, 0.01, 1, "gelu"),
    ((1, 64, 150), 2, 0.01, 1, "silu"),
], ids=["2d_tiles", "1d_block12", "1d_frac", "3d_odd", "2d_bs1", "2d_block64", "1d_block32"])
def test_emulated_afno_matches_float64_reference_algorithm(case):
    shape, nb, lam, frac, act = case
    rng = np.random.default_rng(11)
    Cn = shape[1]
    bs = Cn // nb
    x = rng.standard_normal(shape).astype(np.float32)
    p = [((rng.standard_normal(s) + 1j * rng.standard_normal(s)) * sc).astype(np.complex64)
         for s, sc in (((nb, bs, bs), 0.7 / bs ** 0.5), ((nb, bs), 0.3), ((nb, bs, bs), 0.7 / bs ** 0.5), ((nb, bs), 0.05))]
    cot = rng.standard_normal(shape).astype(np.float32)
    kept = fno_port.afno_kept_modes(shape[2:], frac)
    layer = emu_harness.EmuAfno(x, *p, nb, kept, act, lam)
    y = layer.forward()
    gx, gp = layer.backward(cot)
    ry, rgx, rgp = _reference_f64(x, p, nb, lam, frac, act, cot)
    assert rel_l2(y, ry) < TOL
    assert rel_l2(gx, rgx) < TOL
    for got, want, k in zip(gp, rgp, ("weights1", "bias1", "weights2", "bias2")):
        assert rel_l2(got, want) < TOL, k
    # the two halves of the backward on their own
    gx2, _ = layer.backward(cot, want_params=False)
    assert np.array_equal(gx, gx2)
    _, gp2 = layer.backward(cot, want_gx=False)
    for a, b in zip(gp, gp2):
        assert np.array_equal(a, b)


def test_emulated_afno_rejects_what_the_reference_rejects():
    from deno4pytorch_b200 import _capi
    x = np.zeros((1, 6, 8), np.float32)
    z2, z1 = np.zeros((4, 1, 1), np.complex64), np.zeros((4, 1), np.complex64)
    with pytest.raises(_capi.DenoError, match="divisible"):
        emu_harness.EmuAfno(x, z2, z1, z2, z1, 4, 5, "relu", 0.01)
    z2, z1 = np.zeros((2, 3, 3), np.complex64), np.zeros((2, 3), np.complex64)
    with pytest.raises(_capi.DenoError, match="modes_last"):
        emu_harness.EmuAfno(x, z2, z1, z2, z1, 2, 6, "relu", 0.01)


def test_emulated_afno_limits_fail_loudly():
    """Outside what the kernels cover the call returns an error - never a silent fallback."""
    from deno4pytorch_b200 import _capi
    z = lambda *s: np.zeros(s, np.complex64)
    with pytest.raises(_capi.DenoError, match="above 64"):                  # block size 128
        emu_harness.EmuAfno(np.zeros((1, 128, 8), np.float32), z(1, 128, 128), z(1, 128), z(1, 128, 128), z(1, 128), 1, 5, "gelu", 0.01)
    layer = emu_harness.EmuAfno(np.zeros((1, 4, 8, 260), np.float32), z(1, 4, 4), z(1, 4), z(1, 4, 4), z(1, 4), 1, 131, "gelu", 0.01)
    with pytest.raises(_capi.DenoError, match="last axis"):                 # 131 bins: beyond the plane kernel's thread layout
        layer.forward()


def test_afno_host_queries_through_the_built_library():
    """The shape bookkeeping entry points of include/deno_afno.h are pure host code: check them on the nvcc-built .so."""
    import ctypes as C
    from deno4pytorch_b200 import _capi, build
    lib = _capi.bind(C.CDLL(build.build()))
    d = _capi.make_afno_desc(10, 64, 4, (55, 64), 33, "gelu", 0.01)
    assert lib.deno_afno_num_modes(C.byref(d)) == 55 * 33                    # every row frequency, 33 half-spectrum bins
    nb = lib.deno_afno_twiddle_bytes(C.byref(d))
    tw = np.zeros(nb // 4, np.float32)
    assert lib.deno_afno_twiddle_fill(C.byref(d), tw.ctypes.data_as(C.c_void_p), nb) == 0
    assert np.array_equal(tw[-64 * 64:].reshape(64, 64), np.eye(64, dtype=np.float32))
    spec = 10 * 64 * 55 * 33 * 8
    assert lib.deno_afno_workspace_bytes(C.byref(d), 0) >= 2 * spec
    assert lib.deno_afno_workspace_bytes(C.byref(d), 1) >= 2 * spec
    d3 = _capi.make_afno_desc(2, 8, 2, (5, 4, 6), 4, "relu", 0.0)
    assert lib.deno_afno_num_modes(C.byref(d3)) == 5 * 4 * 4
    bad = _capi.make_afno_desc(1, 6, 4, (8,), 5, "relu", 0.01)
    assert lib.deno_afno_twiddle_bytes(C.byref(bad)) == 0 and b"divisible" in lib.deno_last_error()
    bad = _capi.make_afno_desc(1, 8, 4, (8,), 5, "relu", -1.0)
    assert lib.deno_afno_num_modes(C.byref(bad)) == 0 and b"sparsity_threshold" in lib.deno_last_error()
