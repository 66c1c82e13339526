"""Seeded random sweep of small / odd / degenerate shapes through the fiber-emulated C ABI (CPU): every layer family -
1-D / 2-D / 3-D spectral convolution, the 4-D stand-alone layer, the adaptive Fourier layers - forward and EVERY gradient
against the reference algorithm (oracle/fno_port.py) evaluated in float64.  Covers what the fixed cases do not: sizes of 1
and 2, modes at every bound (all bins of the half spectrum, 2 * modes == N on a leading axis), one channel, block sizes that
are not multiples of 4, ragged last tiles.  Tolerance 1e-5 relative L2 (north_star); a gradient that is a SUM over all points
of a single channel (bypass bias / weight with one output channel) is ill-conditioned when the sum cancels, so those are also
accepted within the fp32 summation bound 2e-6 * sum |terms|."""
import numpy as np
import pytest
import torch

import emu_harness
from golden_util import rel_l2
from oracle import fno_port

TOL = 1e-5
ACTS = ["identity", "gelu", "silu", "relu", "tanh", "leakyrelu"]


def _close(got, want, slack=0.0):
    got, want = torch.as_tensor(got), torch.as_tensor(want)
    if rel_l2(got, want) < TOL:
        return True
    return slack > 0 and (got.double() - want.double()).abs().max().item() < slack


def _spectral_case(rng):
    nd = int(rng.integers(1, 5))
    if nd == 1:
        n = 2 * int(rng.integers(1, 40))
        spatial, modes = (n,), (int(rng.integers(1, n // 2 + 2)),)
    elif nd == 2:
        h, w = int(rng.integers(2, 20)), int(rng.integers(1, 30))
        spatial, modes = (h, w), (int(rng.integers(1, h // 2 + 1)), int(rng.integers(1, w // 2 + 2)))
    elif nd == 3:
        x, h, w = int(rng.integers(2, 8)), int(rng.integers(2, 9)), int(rng.integers(1, 12))
        spatial = (x, h, w)
        modes = (int(rng.integers(1, x // 2 + 1)), int(rng.integers(1, h // 2 + 1)), int(rng.integers(1, w // 2 + 2)))
    else:
        x, y, z, t = int(rng.integers(2, 6)), int(rng.integers(2, 6)), int(rng.integers(1, 7)), int(rng.integers(1, 6))
        spatial = (x, y, z, t)
        modes = (int(rng.integers(1, x // 2 + 1)), int(rng.integers(1, y // 2 + 1)), int(rng.integers(1, z + 1)),
                 int(rng.integers(1, t // 2 + 2)))
    return nd, spatial, modes


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_fuzz_spectral_layers_1d_to_4d(seed):
    rng = np.random.default_rng(seed)
    for _ in range(60):
        nd, spatial, modes = _spectral_case(rng)
        B, cin, cout = int(rng.integers(1, 4)), int(rng.integers(1, 7)), int(rng.integers(1, 7))
        act = str(rng.choice(ACTS))
        nblk = 4 if nd == 4 else 2 ** (nd - 1)
        cplx = lambda s: rng.standard_normal(s) + 1j * rng.standard_normal(s)
        x = torch.from_numpy(rng.standard_normal((B, cin) + spatial).astype(np.float32))
        ws = [torch.from_numpy((cplx((cin, cout) + modes) / cin).astype(np.complex64)) for _ in range(nblk)]
        lw = torch.from_numpy((rng.standard_normal((cout, cin)) / cin ** 0.5).astype(np.float32))
        lb = torch.from_numpy(rng.standard_normal(cout).astype(np.float32))
        cot = torch.from_numpy(rng.standard_normal((B, cout) + spatial).astype(np.float32))
        layer = emu_harness.EmuLayer(x.numpy(), [w.numpy() for w in ws], lw.numpy(), lb.numpy(), modes, act)
        y = layer.forward()
        gx, gws, glw, glb = layer.backward(cot.numpy())

        xr = x.double().requires_grad_(True)
        wr = [w.to(torch.complex128).requires_grad_(True) for w in ws]
        lwr, lbr = lw.double().requires_grad_(True), lb.double().requires_grad_(True)
        if nd == 4:
            spec = fno_port.spectral_conv4d(xr, wr, modes)
            byp = torch.nn.functional.conv1d(xr.reshape(B, cin, -1), lwr.reshape(cout, cin, 1), lbr).reshape(spec.shape)
            z = spec + byp
            ref = fno_port.activation_fn(act)(z)
        else:
            ref = fno_port.spectral_conv(xr, wr, lwr.reshape((cout, cin) + (1,) * nd), lbr, modes, activation=act, norm=None)
        ref.backward(cot.double())
        tag = (spatial, modes, (B, cin, cout), act)
        assert _close(y, ref.detach()), tag
        assert _close(gx, xr.grad), tag
        for a, b in zip(gws, wr):
            assert _close(a, b.grad), tag
        sum_bound = 2e-6 * cot.abs().sum().item()            # |gz| <= |gy| * max|act'| (<= 1.13 for GELU)
        assert _close(glw.reshape(cout, cin), lwr.grad, sum_bound * 1.2 * x.abs().max().item()), tag
        assert _close(glb, lbr.grad, sum_bound * 1.2), tag


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_fuzz_adaptive_fourier_layers(seed):
    rng = np.random.default_rng(seed)
    done = 0
    while done < 80:
        nd = int(rng.integers(1, 4))
        spatial = [(int(rng.integers(1, 80)),), (int(rng.integers(1, 20)), int(rng.integers(1, 40))),
                   (int(rng.integers(1, 7)), int(rng.integers(1, 8)), int(rng.integers(1, 14)))][nd - 1]
        nb, bs = int(rng.choice([1, 2, 3, 4])), int(rng.choice([1, 2, 3, 5, 8, 9]))
        B = int(rng.integers(1, 4))
        frac, lam = float(rng.choice([1, 1, 0.5, 0.2])), float(rng.choice([0.0, 0.01, 0.1]))
        act = str(rng.choice(ACTS))
        kept = fno_port.afno_kept_modes(spatial, frac)
        if kept == 0:
            continue
        done += 1
        shape = (B, nb * bs) + spatial
        cplx = lambda s, sc: ((rng.standard_normal(s) + 1j * rng.standard_normal(s)) * sc).astype(np.complex64)
        x = rng.standard_normal(shape).astype(np.float32)
        p = [cplx((nb, bs, bs), 0.7 / bs ** 0.5), cplx((nb, bs), 0.3), cplx((nb, bs, bs), 0.7 / bs ** 0.5), cplx((nb, bs), 0.05)]
        cot = rng.standard_normal(shape).astype(np.float32)
        layer = emu_harness.EmuAfno(x, *p, nb, kept, act, lam)
        y = layer.forward()
        gx, gp = layer.backward(cot)
        xr = torch.from_numpy(x).double().requires_grad_(True)
        pr = [torch.from_numpy(a).to(torch.complex128).requires_grad_(True) for a in p]
        ref = fno_port.afno_filter(xr, *pr, num_blocks=nb, sparsity_threshold=lam, hard_thresholding_fraction=frac,
                                   activation=act)
        ref.backward(torch.from_numpy(cot).double())
        tag = (shape, nb, kept, act, lam)
        assert _close(y, ref.detach()), tag
        assert _close(gx, xr.grad), tag
        for a, b in zip(gp, pr):
            assert _close(a, b.grad, 1e-7 if b.grad.abs().max() == 0 else 0.0), tag     # fully shrunk: exact zeros


@pytest.mark.parametrize("seed", [1, 2])
def test_fuzz_lift_and_projection(seed):
    """include/deno_fno.h over random grids (1 .. 300 points per axis, any padding, every built width, 1 .. 4 outputs) against
    the torch-op formulation of FNOs.py:136-152 in float64."""
    import torch.nn.functional as F
    rng = np.random.default_rng(seed)
    f32 = lambda t: t.detach().numpy().astype(np.float32)
    for _ in range(40):
        nd = int(rng.integers(1, 4))
        spatial = [(int(rng.integers(1, 300)),), (int(rng.integers(1, 24)), int(rng.integers(1, 24))),
                   (int(rng.integers(1, 7)), int(rng.integers(1, 7)), int(rng.integers(1, 9)))][nd - 1]
        pad = int(rng.choice([0, 0, 1, 2, 5, 9]))
        Cw, Hd = int(rng.choice([8, 12, 16, 20, 24, 32, 40, 48, 64])), int(rng.choice([4, 16, 24, 32, 128]))
        Kx, O, B = int(rng.integers(1, 17 - nd)), int(rng.integers(1, 5)), int(rng.integers(1, 4))
        e = emu_harness.EmuEnds(B, Cw, Hd, Kx, nd, O, spatial, pad)
        assert e.supported()
        dd = torch.float64
        g = torch.Generator().manual_seed(int(rng.integers(0, 1 << 30)))
        rn = lambda *s: torch.randn(*s, generator=g, dtype=dd)
        x = rn((B,) + spatial + (Kx,)).requires_grad_(True)
        grid = torch.rand((B,) + spatial + (nd,), generator=g, dtype=dd)
        w0, b0 = (rn(Cw, Kx + nd) * 0.5).requires_grad_(True), (rn(Cw) * 0.1).requires_grad_(True)
        w1, b1 = (rn(Hd, Cw) / Cw ** 0.5).requires_grad_(True), (rn(Hd) * 0.1).requires_grad_(True)
        w2, b2 = (rn(O, Hd) / Hd ** 0.5).requires_grad_(True), (rn(O) * 0.1).requires_grad_(True)
        h = F.linear(torch.cat((x, grid), -1), w0, b0).movedim(-1, 1)
        if pad:
            h = F.pad(h, [0, pad] * nd)
        gh = rn(h.shape)
        h.backward(gh)
        tag = (B, Cw, Hd, Kx, O, spatial, pad)
        assert _close(e.lift(f32(x), f32(grid), f32(w0), f32(b0)), h.detach()), tag
        gx, gw0, gb0 = e.lift_bwd(f32(gh), f32(x), f32(grid), f32(w0))
        assert _close(gx, x.grad) and _close(gw0, w0.grad) and _close(gb0, b0.grad, 2e-6 * gh.abs().sum().item()), tag
        hp = rn((B, Cw) + tuple(s + pad for s in spatial)).requires_grad_(True)
        crop = hp[(slice(None), slice(None)) + tuple(slice(0, s) for s in spatial)]
        y = F.linear(F.gelu(F.linear(crop.movedim(1, -1), w1, b1)), w2, b2)
        gy = rn(y.shape)
        y.backward(gy)
        assert _close(e.proj(f32(hp), f32(w1), f32(b1), f32(w2), f32(b2)), y.detach()), tag
        ghp, gw1, gb1, gw2, gb2 = e.proj_bwd(f32(gy), f32(hp), f32(w1), f32(b1), f32(w2))
        assert _close(ghp, hp.grad) and _close(gw1, w1.grad) and _close(gb1, b1.grad) and _close(gw2, w2.grad), tag
        assert _close(gb2, b2.grad, 2e-6 * gy.abs().sum().item()), tag        # one output: a plain sum that may cancel
