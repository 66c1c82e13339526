"""Torch-CPU port of the reference FNO spectral path (TEST INFRASTRUCTURE ONLY).

Restates, dimension-generically, what the reference does with library calls:

* layer   : /root/reference/Models/fno/spectral_layers.py:84-115 (1-D),
            :183-208 (2-D), :299-338 (3-D)
* mixing  : spectral_layers.py:73-82, :172-181, :287-297 (complex einsum, or the
            four real einsums of the ``(..., 2)`` real-view weight layout)
* stacks  : /root/reference/Models/fno/FNOs.py:61-83, :131-152, :197-219
* act map : /root/reference/Models/configs.py:17-19

Parity status: PINNED - ``tests/test_oracle_golden.py`` checks this port
against fixtures generated from the real reference by
``tests/golden/make_golden.py`` (run in the build container where
``/root/reference`` is mounted), and, when that mount is present, against the
imported reference directly.

Kept deliberately FFT-based (not truncated-DFT-based) so it is an independent
check on ``oracle/dft_truth.py`` and on the CUDA kernels, and so that timing it
is a fair stand-in for the reference's CPU path (``bench.py --impl reference``).
"""
from __future__ import annotations

import itertools
import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

_LETTERS = "xyz"


def activation_fn(name):
    """Reference activation table, configs.py:17-19 (``None`` maps to SiLU)."""
    table = {
        "gelu": lambda t: F.gelu(t),                     # erf form (nn.GELU() default)
        "silu": F.silu,
        "relu": F.relu,
        "tanh": torch.tanh,
        "leakyrelu": lambda t: F.leaky_relu(t, 0.01),
        None: F.silu,
        "identity": lambda t: t,
    }
    return table[name]


def normalize_modes(modes, ndim: int) -> Tuple[int, ...]:
    """int -> same count on each axis (spectral_layers.py:143-148, 237-244)."""
    if isinstance(modes, int):
        return (modes,) * ndim
    modes = tuple(int(m) for m in modes)
    assert len(modes) == ndim
    return modes


def corner_slices(modes: Sequence[int]) -> List[Tuple[slice, ...]]:
    """Frequency-domain corner blocks in the reference's weight order.

    1-D: [:m]                                   (spectral_layers.py:102)
    2-D: [:m1,:m2], [-m1:,:m2]                  (spectral_layers.py:196-199)
    3-D: (+,+), (-,+), (+,-), (-,-) on (x, y)   (spectral_layers.py:318-325)
    The last (half-spectrum) axis always keeps the low block only.
    """
    lead = modes[:-1]
    last = slice(0, modes[-1])
    out = []
    # the reference toggles the sign of the FIRST axis fastest
    for signs in itertools.product((0, 1), repeat=len(lead)):
        signs = signs[::-1]
        sl = tuple(slice(0, m) if s == 0 else slice(-m, None) for m, s in zip(lead, signs))
        out.append(sl + (last,))
    return out


def _as_complex(w: torch.Tensor) -> torch.Tensor:
    return w if w.is_complex() else torch.view_as_complex(w.contiguous())


def spectral_conv(x: torch.Tensor,
                  weights: Sequence[torch.Tensor],
                  lin_w: torch.Tensor,
                  lin_b: Optional[torch.Tensor],
                  modes,
                  activation="gelu",
                  norm=None,
                  return_freq: bool = False):
    """One spectral layer, any of 1/2/3 spatial dims, channels-first input.

    ``weights`` are ``(Cin, Cout, *modes)`` complex64 or ``(Cin, Cout, *modes, 2)``
    float32 tensors, 1/2/4 of them in the reference's block order.
    """
    ndim = x.dim() - 2
    modes = normalize_modes(modes, ndim)
    dims = list(range(-ndim, 0))
    spatial = x.shape[2:]
    letters = _LETTERS[:ndim]

    bypass = F.conv1d(x.flatten(2), lin_w.reshape(lin_w.shape[0], lin_w.shape[1], 1), lin_b)
    bypass = bypass.reshape(x.shape[0], lin_w.shape[0], *spatial)

    x_ft = torch.fft.rfftn(x, dim=dims, norm=norm)
    out_shape = (x.shape[0], lin_w.shape[0]) + tuple(spatial[:-1]) + (spatial[-1] // 2 + 1,)
    out_ft = torch.zeros(out_shape, dtype=x_ft.dtype, device=x.device)
    eq = f"bi{letters},io{letters}->bo{letters}"
    for w, sl in zip(weights, corner_slices(modes)):
        idx = (slice(None), slice(None)) + sl
        out_ft[idx] = torch.einsum(eq, x_ft[idx], _as_complex(w))

    if ndim == 1:
        # spectral_layers.py:106 calls irfft WITHOUT n= : length becomes 2*(N//2)
        spec = torch.fft.irfft(out_ft, norm=norm)
    else:
        spec = torch.fft.irfftn(out_ft, s=tuple(spatial), dim=dims, norm=norm)
    y = activation_fn(activation)(spec + bypass)
    if return_freq:
        return y, out_ft
    return y


def _pad_last(x: torch.Tensor, ndim: int, padding: int) -> torch.Tensor:
    if padding == 0:
        return x
    pads = []
    for _ in range(ndim):
        pads += [0, padding]
    return F.pad(x, pads)


def fno_forward(params: Dict[str, torch.Tensor], x: torch.Tensor, grid: torch.Tensor, *,
                ndim: int, modes, depth: int, padding: int, activation="gelu") -> torch.Tensor:
    """FNO1d/2d/3d forward with a reference-format state dict (FNOs.py:61-83,131-152,197-219).

    ``params`` uses the reference's state-dict keys: ``fc0.weight``, ``convs.{i}.weights{j}``,
    ``convs.{i}.linear.weight`` ...
    """
    modes = normalize_modes(modes, ndim)
    nblk = 2 ** (ndim - 1)
    h = torch.cat((x, grid), dim=-1)
    h = F.linear(h, params["fc0.weight"], params["fc0.bias"])
    h = h.movedim(-1, 1)
    h = _pad_last(h, ndim, padding)
    for i in range(depth):
        ws = [params[f"convs.{i}.weights{j + 1}"] for j in range(nblk)]
        h = spectral_conv(h, ws, params[f"convs.{i}.linear.weight"], params[f"convs.{i}.linear.bias"],
                          modes, activation=activation, norm=None)
    if padding != 0:
        crop = (Ellipsis,) + (slice(None, -padding),) * ndim
        h = h[crop]
    h = h.movedim(1, -1)
    h = F.gelu(F.linear(h, params["fc1.weight"], params["fc1.bias"]))
    return F.linear(h, params["fc2.weight"], params["fc2.bias"])


def init_fno_params(ndim: int, in_dim: int, out_dim: int, modes, width: int, depth: int = 4,
                    hidden: int = 128, steps: int = 1, use_complex: bool = True,
                    seed: int = 0) -> Dict[str, torch.Tensor]:
    """Random parameters with the reference's shapes and init distributions.

    ``scale * U[0,1)`` on real and imaginary parts of the spectral weights
    (spectral_layers.py:63-68,157-169,253-284); torch defaults (Kaiming-uniform)
    for Linear / 1x1 conv.  NOT bit-identical to the reference's RNG stream -
    fixtures carry explicit weights instead.
    """
    g = torch.Generator().manual_seed(seed)
    modes = normalize_modes(modes, ndim)
    p: Dict[str, torch.Tensor] = {}

    def uniform(shape, bound):
        return (torch.rand(shape, generator=g) * 2 - 1) * bound

    def linear(name, fan_in, fan_out, extra=()):
        bound = 1.0 / math.sqrt(fan_in)
        p[f"{name}.weight"] = uniform((fan_out, fan_in) + tuple(extra), bound)
        p[f"{name}.bias"] = uniform((fan_out,), bound)

    linear("fc0", steps * in_dim + ndim, width)
    scale = 1.0 / (width * width)
    for i in range(depth):
        for j in range(2 ** (ndim - 1)):
            w = scale * torch.rand((width, width) + modes + (2,), generator=g)
            p[f"convs.{i}.weights{j + 1}"] = torch.view_as_complex(w) if use_complex else w
        linear(f"convs.{i}.linear", width, width, extra=(1,) * ndim)
    linear("fc1", width, hidden)
    linear("fc2", hidden, out_dim)
    return p


def make_grid(batch: int, spatial: Sequence[int]) -> torch.Tensor:
    """``linspace(0,1)`` coordinate mesh appended by every demo's ``feature_transform``
    (/root/reference/Demo/Darcy_2d/run_FNO&UNet.py:27-40)."""
    axes = [torch.linspace(0, 1, n, dtype=torch.float32) for n in spatial]
    mesh = torch.stack(torch.meshgrid(*axes, indexing="ij"), dim=-1)
    return mesh.unsqueeze(0).expand(batch, *mesh.shape).contiguous()


# ----------------------------------------------------------------------------------------------------------------
# 4-D stand-alone layer / model (not on the CUDA path yet: this restatement + its fixtures
# pin the oracle first).
# ----------------------------------------------------------------------------------------------------------------
def spectral_conv4d(x: torch.Tensor, weights: Sequence[torch.Tensor], modes) -> torch.Tensor:
    """``SpectralConv4d.forward`` (/root/reference/Demo/Turbulence_3d+t/FNO_HIT.py:31-78): ``rfftn`` over the last four
    axes, FOUR corner blocks - both signs on the first two axes only, the LOW block ``[:m3]`` on the third axis (its
    negative frequencies are dropped, unlike the 3-D layer) and ``[:m4]`` on the half-spectrum axis - one complex
    einsum per block, zero-padded ``irfftn(s=...)``.  No bypass, bias or activation inside the layer.  The reference
    clamps ``modes4`` to ``min(modes4, 3 // 2 + 1) = 2`` in the constructor (FNO_HIT.py:44): pass the clamped value."""
    m1, m2, m3, m4 = (int(m) for m in modes)
    dims = [-4, -3, -2, -1]
    spatial = tuple(x.shape[2:])
    x_ft = torch.fft.rfftn(x, dim=dims)
    cout = weights[0].shape[1]
    out_ft = torch.zeros((x.shape[0], cout) + spatial[:-1] + (spatial[-1] // 2 + 1,), dtype=x_ft.dtype, device=x.device)
    low3, low4 = slice(0, m3), slice(0, m4)
    blocks = [(slice(0, m1), slice(0, m2)), (slice(-m1, None), slice(0, m2)),
              (slice(0, m1), slice(-m2, None)), (slice(-m1, None), slice(-m2, None))]      # weights1..4 (FNO_HIT.py:63-72)
    for w, (s1, s2) in zip(weights, blocks):
        idx = (slice(None), slice(None), s1, s2, low3, low4)
        out_ft[idx] = torch.einsum("bixyzt,ioxyzt->boxyzt", x_ft[idx], _as_complex(w))
    return torch.fft.irfftn(out_ft, s=spatial, dim=dims)


def make_grid4d(batch: int, spatial: Sequence[int]) -> torch.Tensor:
    """``FNO4d.get_grid`` (FNO_HIT.py:159-169): linspace(0, 1) coordinates of the four axes, channel-last."""
    axes = [torch.linspace(0, 1, n, dtype=torch.float32) for n in spatial]
    mesh = torch.stack(torch.meshgrid(*axes, indexing="ij"), dim=-1)
    return mesh.unsqueeze(0).expand(batch, *mesh.shape).contiguous()


def fno4d_forward(params: Dict[str, torch.Tensor], x: torch.Tensor, modes) -> torch.Tensor:
    """``FNO4d.forward`` (FNO_HIT.py:118-157): ``cat(x, grid) -> fc0 -> 4 x (SpectralConv4d + 1x1 conv), GELU after the
    first three only -> fc1 -> GELU -> fc2``; ``x`` is ``(B, X, Y, Z, T, 5)``, the grid adds 4 coordinates (9 features)."""
    B = x.shape[0]
    spatial = tuple(x.shape[1:5])
    h = torch.cat((x, make_grid4d(B, spatial).to(x)), dim=-1)
    h = F.linear(h, params["fc0.weight"], params["fc0.bias"]).permute(0, 5, 1, 2, 3, 4)
    width = h.shape[1]
    for i in range(4):
        ws = [params[f"conv{i}.weights{j}"] for j in (1, 2, 3, 4)]
        x1 = spectral_conv4d(h, ws, modes)
        x2 = F.conv1d(h.reshape(B, width, -1), params[f"w{i}.weight"], params[f"w{i}.bias"]).reshape(B, width, *spatial)
        h = x1 + x2
        if i < 3:
            h = F.gelu(h)
    h = h.permute(0, 2, 3, 4, 5, 1)
    h = F.gelu(F.linear(h, params["fc1.weight"], params["fc1.bias"]))
    return F.linear(h, params["fc2.weight"], params["fc2.bias"])


class PortTrainer:
    """fwd + MSE + bwd + Adam on CPU, the way every FNO demo trains
    (/root/reference/Demo/Darcy_2d/run_FNO&UNet.py:43-69,206-209).  Used as the
    timed CPU baseline."""

    def __init__(self, params: Dict[str, torch.Tensor], *, ndim, modes, depth, padding,
                 activation="gelu", lr=1e-3, betas=(0.7, 0.9), weight_decay=1e-4):
        self.params = {k: v.clone().requires_grad_(True) for k, v in params.items()}
        self.cfg = dict(ndim=ndim, modes=modes, depth=depth, padding=padding, activation=activation)
        self.opt = torch.optim.Adam(list(self.params.values()), lr=lr, betas=betas,
                                    weight_decay=weight_decay)

    def step(self, x, grid, target) -> float:
        pred = fno_forward(self.params, x, grid, **self.cfg)
        loss = F.mse_loss(pred, target)
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()
        return loss.item()


# ----------------------------------------------------------------------------------------------------------------
# Adaptive Fourier layers: AdaptiveFourier1d/2d/3d,
# /root/reference/Models/fno/spectral_layers.py:341-573.  Pinned by tests/golden/afno{1,2,3}d*.npz
# (tests/golden/make_golden_afno.py imports the reference classes).
# ----------------------------------------------------------------------------------------------------------------
def afno_kept_modes(spatial: Sequence[int], hard_thresholding_fraction: float) -> int:
    """How many bins of the HALF-SPECTRUM (last) axis the block-diagonal MLP touches.  The reference computes
    ``total_modes = prod(spatial) // 2 + 1`` (``N // 2 + 1`` in 1-D; spectral_layers.py:394, :470, :547),
    ``kept = int(total_modes * fraction)`` and slices the LAST axis with ``[..., :kept]`` - so in 2-D / 3-D the default
    fraction 1 keeps the whole axis and only a small fraction truncates it.  Bins beyond it stay zero."""
    n = 1
    for v in spatial:
        n *= int(v)
    kept = int((n // 2 + 1) * hard_thresholding_fraction)
    return max(0, min(kept, int(spatial[-1]) // 2 + 1))


def afno_filter(x: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor, w2: torch.Tensor, b2: torch.Tensor, *,
                num_blocks: int, sparsity_threshold: float = 0.01, hard_thresholding_fraction: float = 1.0,
                activation="gelu") -> torch.Tensor:
    """``AdaptiveFourier{1,2,3}d.forward`` (spectral_layers.py:383-409, :457-487, :533-565), any of 1-3 spatial axes:
    ortho ``rfftn`` -> channels split into ``num_blocks`` blocks -> per bin ``act(x W1 + b1) W2 + b2`` with complex
    block-diagonal weights ``(blocks, block, block)``, the activation applied to real and imaginary parts separately
    (``view_as_real``) -> soft-shrink of real and imaginary parts -> ortho ``irfftn(s=...)`` -> ``+ x``.  Only
    ``hidden_size_factor = 1`` runs in the reference (``weights2`` is created ``(blocks, block, block * factor)`` but
    contracted over its FIRST block axis, :372-374/:381)."""
    nd = x.dim() - 2
    dims = list(range(-nd, 0))
    B, C = x.shape[:2]
    spatial = tuple(x.shape[2:])
    act = activation_fn(activation)
    xf = torch.fft.rfftn(x, dim=dims, norm="ortho")
    fshape = tuple(xf.shape[2:])
    bs = C // num_blocks
    xf = xf.reshape((B, num_blocks, bs) + fshape)
    kept = afno_kept_modes(spatial, hard_thresholding_fraction)
    letters = "jkl"[:nd]
    eq = f"bni{letters},nio->bno{letters}"
    tail = (None,) * nd
    cdt = xf.dtype
    h_in = xf[..., :kept]
    h = torch.einsum(eq, h_in, w1.to(cdt)) + b1.to(cdt)[(None, slice(None), slice(None)) + tail]
    h = torch.view_as_complex(act(torch.view_as_real(h)))
    o = torch.einsum(eq, h, w2.to(cdt)) + b2.to(cdt)[(None, slice(None), slice(None)) + tail]
    out = torch.zeros_like(xf)
    out[..., :kept] = o
    out = torch.view_as_complex(F.softshrink(torch.view_as_real(out), lambd=sparsity_threshold))
    out = out.reshape((B, C) + fshape)
    return torch.fft.irfftn(out, s=spatial, dim=dims, norm="ortho") + x
