"""Drop-in ``FNO1d/2d/3d`` stacks over the CUDA spectral layers.

Constructor signatures, attribute names, sub-module names (``fc0``, ``convs``, ``fc1``, ``fc2``)
and ``forward(x, grid)`` follow /root/reference/Models/fno/FNOs.py:26-27,61 (1-D), :91-92,131
(2-D), :160-161,197 (3-D), so state dicts are interchangeable with the reference.
"""
from __future__ import annotations

from .spectral_layers import *  # noqa: F401,F403  (the reference module star-exports these names)
from .spectral_layers import SpectralConv1d, SpectralConv2d, SpectralConv3d, nn, torch, F

_LAYER = {1: SpectralConv1d, 2: SpectralConv2d, 3: SpectralConv3d}


class _TallLinearFn(torch.autograd.Function):
    """``F.linear`` for inputs with very many rows (every grid point of every sample) and few features.

    Forward and grad_input are plain cuBLAS GEMMs.  grad_weight is a GEMM whose reduction dimension is
    the row count (144 500 for the Darcy config); cuBLAS' fp32 heuristics run it as two CTAs looping
    over all rows (239 us per call on B200, profiles/r1i_launch_list.txt), so it is expressed here as a
    batched GEMM over row chunks followed by a short sum (split-K), and grad_bias as a two-stage sum.
    Same arithmetic, different summation order (fp32 both ways)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return F.linear(x, weight, bias)

    @staticmethod
    def backward(ctx, gy):
        x, weight = ctx.saved_tensors
        gx = gw = gb = None
        g2 = gy.reshape(-1, gy.shape[-1])
        if ctx.needs_input_grad[0]:
            gx = (g2 @ weight).view(x.shape)
        rows = g2.shape[0]
        chunks = _split_rows(rows)
        if ctx.needs_input_grad[1]:
            x2 = x.reshape(-1, x.shape[-1])
            if chunks > 1:
                gw = torch.bmm(g2.view(chunks, rows // chunks, -1).transpose(1, 2),
                               x2.view(chunks, rows // chunks, -1)).sum(0)
            else:
                gw = g2.t() @ x2
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = g2.view(chunks, rows // chunks, -1).sum(1).sum(0) if chunks > 1 else g2.sum(0)
        return gx, gw, gb


def _split_rows(rows: int) -> int:
    """Largest divisor of ``rows`` that leaves chunks of at least 512 rows, capped at 512 chunks."""
    if rows < 8192:
        return 1
    best = 1
    for c in range(2, 513):
        if rows % c == 0 and rows // c >= 512:
            best = c
    return best


def _tall_linear(x, layer: nn.Linear):
    if x.is_cuda and x.dtype == torch.float32:
        return _TallLinearFn.apply(x, layer.weight, layer.bias)
    return layer(x)


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
        h = _tall_linear(torch.cat((x, grid), dim=-1), self.fc0)
        h = h.movedim(-1, 1)
        if self.padding != 0:
            h = F.pad(h, [0, self.padding] * nd)
        for layer in self.convs:
            h = layer(h)
        if self.padding != 0:
            h = h[(Ellipsis,) + (slice(None, -self.padding),) * nd]
        h = h.movedim(1, -1)
        return _tall_linear(F.gelu(_tall_linear(h, self.fc1)), self.fc2)


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
