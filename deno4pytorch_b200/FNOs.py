"""Drop-in ``FNO1d/2d/3d`` stacks over the CUDA spectral layers.

Constructor signatures, attribute names, sub-module names (``fc0``, ``convs``, ``fc1``, ``fc2``)
and ``forward(x, grid)`` follow /root/reference/Models/fno/FNOs.py:26-27,61 (1-D), :91-92,131
(2-D), :160-161,197 (3-D), so state dicts are interchangeable with the reference.
"""
from __future__ import annotations

from .spectral_layers import *  # noqa: F401,F403  (the reference module star-exports these names)
from .spectral_layers import SpectralConv1d, SpectralConv2d, SpectralConv3d, nn, torch, F

_LAYER = {1: SpectralConv1d, 2: SpectralConv2d, 3: SpectralConv3d}


class _FNONd(nn.Module):
    ndim = 0

    def _build(self, in_dim, out_dim, modes, width, depth, hidden, steps, padding, activation, use_complex,
               layer_kwargs):
        self.in_dim = in_dim
        self.out_dim = out_dim
        self.modes = modes
        self.width = width
        self.depth = depth
        self.hidden = hidden
        self.steps = steps
        self.activation = activation
        self.use_complex = use_complex
        self.padding = padding  # right-pad of every spatial axis for non-periodic inputs
        self.fc0 = nn.Linear(steps * in_dim + self.ndim, width)  # lift: (history, coordinates) -> width
        self.convs = nn.ModuleList(
            _LAYER[self.ndim](width, width, modes, activation=activation, norm=None, use_complex=use_complex,
                              **layer_kwargs)
            for _ in range(depth))
        self.fc1 = nn.Linear(width, hidden)
        self.fc2 = nn.Linear(hidden, out_dim)

    def forward(self, x, grid):
        """x: (B, *spatial, steps*in_dim), grid: (B, *spatial, ndim) -> (B, *spatial, out_dim)."""
        nd = self.ndim
        h = self.fc0(torch.cat((x, grid), dim=-1))
        h = h.movedim(-1, 1)
        if self.padding != 0:
            h = F.pad(h, [0, self.padding] * nd)
        for layer in self.convs:
            h = layer(h)
        if self.padding != 0:
            h = h[(Ellipsis,) + (slice(None, -self.padding),) * nd]
        h = h.movedim(1, -1)
        return self.fc2(F.gelu(self.fc1(h)))


class FNO1d(_FNONd):
    ndim = 1

    def __init__(self, in_dim, out_dim, modes=16, width=64, depth=4, hidden=128, steps=1, padding=2,
                 activation="gelu", use_complex=True):
        super().__init__()
        self._build(in_dim, out_dim, modes, width, depth, hidden, steps, padding, activation, use_complex, {})


class FNO2d(_FNONd):
    ndim = 2

    def __init__(self, in_dim, out_dim, modes=(8, 8), width=32, depth=4, hidden=128, steps=1, padding=2,
                 activation="gelu", dropout=0.0, use_complex=True):
        super().__init__()
        self.dropout = dropout
        self._build(in_dim, out_dim, modes, width, depth, hidden, steps, padding, activation, use_complex,
                    {"dropout": dropout})


class FNO3d(_FNONd):
    ndim = 3

    def __init__(self, in_dim, out_dim, modes=(8, 8, 8), width=32, depth=4, hidden=128, steps=1, padding=6,
                 activation="gelu", use_complex=True):
        super().__init__()
        self._build(in_dim, out_dim, modes, width, depth, hidden, steps, padding, activation, use_complex, {})
