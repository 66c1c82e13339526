"""Drop-in ``SpectralConv1d/2d/3d`` running on the sm_100a kernels.

Same constructor signatures, defaults, attribute names, parameter names / shapes / dtypes and
``forward(x)`` contract as /root/reference/Models/fno/spectral_layers.py:38-46 (1-D),
:126-134 (2-D), :219-226 (3-D), so reference checkpoints load unchanged and the Demo scripts
run as they are.  What differs is the execution: one fused call into libdeno_spectral.so
(deno4pytorch_b200.functional.spectral_conv) instead of conv + rfftn + einsum + irfftn + add +
activation, with a hand-written backward.

Deliberate differences from the reference (both are reference defects):
  * 2-D layers also work with ``use_complex=False`` (the reference's real-view branch crashes);
  * 3-D layers also work with ``use_complex=True`` (the reference builds ``weights4`` in the wrong
    layout and crashes).  Parameter names and shapes for the working variants are identical.
"""
from __future__ import annotations

import math  # noqa: F401  (re-exported: reference callers star-import this module)
import os    # noqa: F401
import sys   # noqa: F401
from functools import partial  # noqa: F401

import torch
import torch.fft as fft  # noqa: F401
import torch.nn as nn
import torch.nn.functional as F  # noqa: F401
from torch.nn.parameter import Parameter  # noqa: F401

from .configs import activation_dict, activation_kernel_id, default  # noqa: F401
from .functional import afno_filter, spectral_conv

_CONV = {1: nn.Conv1d, 2: nn.Conv2d, 3: nn.Conv3d}


class _SpectralConvNd(nn.Module):
    """Dimension-generic implementation; the public classes only fix ``ndim`` and defaults."""

    ndim = 0
    _dropout_in_forward = False  # only the 2-D reference layer applies its dropout (spectral_layers.py:190)

    def __init__(self, in_dim, out_dim, modes, dropout=0.1, norm="ortho", activation="gelu", return_freq=False,
                 use_complex=True):
        super().__init__()
        nd = self.ndim
        self.in_dim = in_dim
        self.out_dim = out_dim
        mode_tuple = (modes,) * nd if isinstance(modes, int) else tuple(modes)
        if len(mode_tuple) != nd:
            raise ValueError(f"{type(self).__name__} needs {nd} mode counts, got {modes}")
        if nd == 1:
            self.modes = mode_tuple[0]
        else:
            for a, m in enumerate(mode_tuple):
                setattr(self, f"modes{a + 1}", m)
        self.norm = norm
        self.dropout = nn.Dropout(dropout)
        self.return_freq = return_freq
        self.use_complex = use_complex
        self.activation = activation_dict[activation]
        self.linear = _CONV[nd](in_dim, out_dim, 1)  # bypass (residual) 1x1 convolution
        self.scale = 1 / (in_dim * out_dim)
        for j in range(2 ** (nd - 1)):
            shape = (in_dim, out_dim) + mode_tuple
            if use_complex:
                w = self.scale * torch.rand(*shape, dtype=torch.cfloat)
            else:
                w = self.scale * torch.rand(*shape, 2, dtype=torch.float32)
            setattr(self, f"weights{j + 1}", nn.Parameter(w))

    # -- helpers ---------------------------------------------------------------------------
    def _mode_tuple(self):
        if self.ndim == 1:
            return (self.modes,)
        return tuple(getattr(self, f"modes{a + 1}") for a in range(self.ndim))

    def _weights(self):
        return [getattr(self, f"weights{j + 1}") for j in range(2 ** (self.ndim - 1))]

    def _freq_output(self, x):
        """Full zero-padded output spectrum for ``return_freq`` (no caller in the reference uses it;
        built with torch ops from the same parameters, outside the hot path)."""
        nd = self.ndim
        dims = list(range(-nd, 0))
        x_ft = torch.fft.rfftn(x, dim=dims, norm=self.norm)
        out_ft = torch.zeros((x.shape[0], self.out_dim) + tuple(x_ft.shape[2:]), dtype=x_ft.dtype, device=x.device)
        modes = self._mode_tuple()
        letters = "xyz"[:nd]
        for j, w in enumerate(self._weights()):
            sl = []
            for a in range(nd - 1):
                hi = (j >> a) & 1
                sl.append(slice(-modes[a], None) if hi else slice(0, modes[a]))
            sl.append(slice(0, modes[-1]))
            idx = (slice(None), slice(None)) + tuple(sl)
            wc = w if w.is_complex() else torch.view_as_complex(w)
            out_ft[idx] = torch.einsum(f"bi{letters},io{letters}->bo{letters}", x_ft[idx], wc)
        return out_ft if self.use_complex else torch.view_as_real(out_ft)

    # -- forward ---------------------------------------------------------------------------
    def forward(self, x):
        """x: (B, in_dim, *spatial) fp32 CUDA tensor -> (B, out_dim, *spatial)."""
        act_id = activation_kernel_id(self.activation)
        fused_act = act_id if act_id else "identity"
        weights = self._weights()
        if self.return_freq:   # the weights are used a second time below: no direct gradient writes (functional.grad_sink)
            weights = [w.view_as(w) for w in weights]
        lw, lb = self.linear.weight, self.linear.bias
        drop = self._dropout_in_forward and self.training and self.dropout.p > 0
        if not drop:
            y = spectral_conv(x, weights, lw, lb, self._mode_tuple(), fused_act)
            if not act_id:
                y = self.activation(y)
        else:
            # dropout acts on the spectral branch's input only (spectral_layers.py:189-191)
            spec = spectral_conv(self.dropout(x), weights, torch.zeros_like(lw), None, self._mode_tuple(), "identity")
            y = self.activation(spec + self.linear(x))
        if self.return_freq:
            return y, self._freq_output(x)
        return y


class SpectralConv1d(_SpectralConvNd):
    """1-D Fourier layer (reference: spectral_layers.py:31-115)."""
    ndim = 1

    def __init__(self, in_dim, out_dim, modes: int, dropout=0.1, norm="ortho", activation="gelu", return_freq=False,
                 use_complex=True):
        super().__init__(in_dim, out_dim, modes, dropout, norm, activation, return_freq, use_complex)


class SpectralConv2d(_SpectralConvNd):
    """2-D Fourier layer (reference: spectral_layers.py:118-208)."""
    ndim = 2
    _dropout_in_forward = True

    def __init__(self, in_dim, out_dim, modes: tuple, dropout=0.1, norm="ortho", activation="gelu",
                 return_freq=False, use_complex=True):
        super().__init__(in_dim, out_dim, modes, dropout, norm, activation, return_freq, use_complex)


class SpectralConv3d(_SpectralConvNd):
    """3-D Fourier layer (reference: spectral_layers.py:211-338; default activation is SiLU there)."""
    ndim = 3

    def __init__(self, in_dim, out_dim, modes: tuple, dropout=0.1, norm="ortho", activation="silu",
                 return_freq=False, use_complex=True):
        super().__init__(in_dim, out_dim, modes, dropout, norm, activation, return_freq, use_complex)


class _AdaptiveFourierNd(nn.Module):
    """Adaptive Fourier layer (reference: spectral_layers.py:341-573): ortho ``rfftn`` -> per-bin two-layer complex MLP
    with block-diagonal weights, activation on real / imaginary parts -> soft-shrink -> ortho ``irfftn`` -> ``+ x``, as
    ONE fused call (``deno_afno_fwd``, include/deno_afno.h) with a hand-written backward.  Constructor signature,
    attributes, parameter names / shapes / init as the reference's three classes."""
    ndim = 0

    def __init__(self, hidden_size, num_blocks, sparsity_threshold, hard_thresholding_fraction, hidden_size_factor,
                 activation):
        super().__init__()
        assert hidden_size % num_blocks == 0, f"hidden_size {hidden_size} should be divisble by num_blocks {num_blocks}"
        self.hidden_size = hidden_size
        self.sparsity_threshold = sparsity_threshold
        self.num_blocks = num_blocks
        self.block_size = self.hidden_size // self.num_blocks
        self.hard_thresholding_fraction = hard_thresholding_fraction
        self.hidden_size_factor = hidden_size_factor
        self.scale = 0.02
        self.activation = activation_dict[activation]
        nb, bs, f = self.num_blocks, self.block_size, self.hidden_size_factor
        self.weights1 = nn.Parameter(self.scale * torch.rand(nb, bs, bs * f, dtype=torch.cfloat))
        self.weights2 = nn.Parameter(self.scale * torch.rand(nb, bs, bs * f, dtype=torch.cfloat))
        self.bias1 = nn.Parameter(self.scale * torch.rand(nb, bs * f, dtype=torch.cfloat))
        self.bias2 = nn.Parameter(self.scale * torch.rand(nb, bs, dtype=torch.cfloat))

    def forward(self, x):
        """x: (B, hidden_size, *spatial) fp32 CUDA tensor -> same shape."""
        if x.dim() != self.ndim + 2:
            raise RuntimeError(f"expected a (B, C, {self.ndim} spatial dims) input, got {tuple(x.shape)}")
        if self.hidden_size_factor != 1:
            # weights2 (blocks, block, block * factor) is contracted over its FIRST block axis with block * factor hidden
            # channels (spectral_layers.py:372-381): the reference's einsum raises for any factor but 1
            raise RuntimeError("hidden_size_factor != 1 fails in the reference's second einsum (operand sizes do not match)")
        act_id = activation_kernel_id(self.activation)
        if not act_id:
            raise RuntimeError(f"activation {type(self.activation).__name__} has no fused kernel "
                               "(identity / gelu / silu / relu / tanh / leakyrelu(0.01))")
        return afno_filter(x, self.weights1, self.bias1, self.weights2, self.bias2, num_blocks=self.num_blocks,
                           sparsity_threshold=self.sparsity_threshold,
                           hard_thresholding_fraction=self.hard_thresholding_fraction, activation=act_id)


class AdaptiveFourier1d(_AdaptiveFourierNd):
    """reference: spectral_layers.py:341-409 (default activation ReLU)."""
    ndim = 1

    def __init__(self, hidden_size, num_blocks=8, sparsity_threshold=0.01, hard_thresholding_fraction=1,
                 hidden_size_factor=1, activation="relu"):
        super().__init__(hidden_size, num_blocks, sparsity_threshold, hard_thresholding_fraction, hidden_size_factor, activation)


class AdaptiveFourier2d(_AdaptiveFourierNd):
    """reference: spectral_layers.py:412-487."""
    ndim = 2

    def __init__(self, hidden_size, num_blocks=8, sparsity_threshold=0.01, hard_thresholding_fraction=1,
                 hidden_size_factor=1, activation="gelu"):
        super().__init__(hidden_size, num_blocks, sparsity_threshold, hard_thresholding_fraction, hidden_size_factor, activation)


class AdaptiveFourier3d(_AdaptiveFourierNd):
    """reference: spectral_layers.py:490-565."""
    ndim = 3

    def __init__(self, hidden_size, num_blocks=8, sparsity_threshold=0.01, hard_thresholding_fraction=1,
                 hidden_size_factor=1, activation="gelu"):
        super().__init__(hidden_size, num_blocks, sparsity_threshold, hard_thresholding_fraction, hidden_size_factor, activation)
