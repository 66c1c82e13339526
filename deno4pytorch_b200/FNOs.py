"""Drop-in ``FNO1d/2d/3d`` stacks over the CUDA spectral layers.

Constructor signatures, attribute names, sub-module names (``fc0``, ``convs``, ``fc1``, ``fc2``)
and ``forward(x, grid)`` follow /root/reference/Models/fno/FNOs.py:26-27,61 (1-D), :91-92,131
(2-D), :160-161,197 (3-D), so state dicts are interchangeable with the reference.
"""
from __future__ import annotations

import ctypes as _C
import os as _os

from . import _capi
from ._lib import lib as _lib
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


def _ptr(t):
    return None if t is None else t.data_ptr()


def _grad_sink(param, like):
    from .functional import _claim_sink
    return _claim_sink(param, like)


def _resolve_sinks(owners, sinks):
    from .functional import _release_sink, _resolve_sink
    out = tuple(_resolve_sink(o, s) for o, s in zip(owners, sinks))
    for o in owners:
        _release_sink(o)
    return out


def _poison(*tensors):
    from .functional import _poison as poison
    poison(*tensors)


def _count_launches(L):
    from .functional import _LAUNCHES
    _LAUNCHES["count"] += L.deno_spectral_last_launch_count()


class _LiftFn(torch.autograd.Function):
    """``pad(permute(fc0(cat(x, grid))))`` as one kernel (libdeno_spectral's ``deno_fno_lift_fwd/_bwd``,
    include/deno_fno.h; /root/reference/Models/fno/FNOs.py:136-142)."""

    @staticmethod
    def forward(ctx, x, grid, w0_in, b0_in, desc, save=True):
        L = _lib()
        x, grid, w0, b0 = x.contiguous(), grid.contiguous(), w0_in.contiguous(), b0_in.contiguous()
        nd = desc.ndim
        padded = tuple(desc.spatial[a] + desc.padding for a in range(nd))
        h = torch.empty((desc.batch, desc.width) + padded, dtype=torch.float32, device=x.device)
        _poison(h)
        with torch.cuda.device(x.device):
            st = torch.cuda.current_stream(x.device).cuda_stream
            _capi.check(L, L.deno_fno_lift_fwd(_C.byref(desc), x.data_ptr(), grid.data_ptr(), w0.data_ptr(),
                                               b0.data_ptr(), h.data_ptr(), st), "deno_fno_lift_fwd")
        _count_launches(L)
        if not save:          # caller is under torch.no_grad(): nothing to save, no gradient sink to claim
            return h
        ctx.save_for_backward(x, grid, w0)
        ctx.sinks = (_grad_sink(w0_in, w0), _grad_sink(b0_in, b0))
        ctx.sink_owners = (w0_in, b0_in)
        ctx.desc = desc
        return h

    @staticmethod
    def backward(ctx, gh):
        L = _lib()
        x, grid, w0 = ctx.saved_tensors
        desc = ctx.desc
        gh = gh.contiguous()
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        s_w0, s_b0 = _resolve_sinks(ctx.sink_owners, ctx.sinks)
        gw0 = s_w0 if s_w0 is not None else torch.empty_like(w0)
        gb0 = s_b0 if s_b0 is not None else torch.empty(w0.shape[0], dtype=torch.float32, device=gh.device)
        nws = L.deno_fno_workspace_bytes(_C.byref(desc), _capi.FNO_WS_LIFT_BWD)
        ws = torch.empty(max(nws, 1), dtype=torch.uint8, device=gh.device)
        _poison(ws, gx, gw0 if s_w0 is None else None, gb0 if s_b0 is None else None)
        with torch.cuda.device(gh.device):
            st = torch.cuda.current_stream(gh.device).cuda_stream
            _capi.check(L, L.deno_fno_lift_bwd(_C.byref(desc), gh.data_ptr(), x.data_ptr(), grid.data_ptr(),
                                               w0.data_ptr(), _ptr(gx), gw0.data_ptr(), gb0.data_ptr(),
                                               ws.data_ptr(), ws.numel(), st), "deno_fno_lift_bwd")
        _count_launches(L)
        return gx, None, None if s_w0 is not None else gw0, None if s_b0 is not None else gb0, None, None


class _ProjFn(torch.autograd.Function):
    """``fc2(gelu(fc1(permute(crop(h)))))`` as one kernel; the hidden layer is never stored and is recomputed
    in the backward pass (``deno_fno_proj_fwd/_bwd``; FNOs.py:147-152 of the reference)."""

    @staticmethod
    def forward(ctx, h, w1_in, b1_in, w2_in, b2_in, desc, save=True):
        L = _lib()
        h, w1, b1, w2, b2 = (t.contiguous() for t in (h, w1_in, b1_in, w2_in, b2_in))
        nd = desc.ndim
        y = torch.empty((desc.batch,) + tuple(desc.spatial[a] for a in range(nd)) + (desc.out_dim,),
                        dtype=torch.float32, device=h.device)
        _poison(y)
        with torch.cuda.device(h.device):
            st = torch.cuda.current_stream(h.device).cuda_stream
            _capi.check(L, L.deno_fno_proj_fwd(_C.byref(desc), h.data_ptr(), w1.data_ptr(), b1.data_ptr(),
                                               w2.data_ptr(), b2.data_ptr(), y.data_ptr(), st), "deno_fno_proj_fwd")
        _count_launches(L)
        if not save:
            return y
        ctx.save_for_backward(h, w1, b1, w2)
        ctx.sinks = tuple(_grad_sink(a, c) for a, c in ((w1_in, w1), (b1_in, b1), (w2_in, w2), (b2_in, b2)))
        ctx.sink_owners = (w1_in, b1_in, w2_in, b2_in)
        ctx.desc = desc
        return y

    @staticmethod
    def backward(ctx, gy):
        L = _lib()
        h, w1, b1, w2 = ctx.saved_tensors
        desc = ctx.desc
        gy = gy.contiguous()
        dev = gy.device
        gh = torch.empty_like(h)
        sinks = _resolve_sinks(ctx.sink_owners, ctx.sinks)
        gw1 = sinks[0] if sinks[0] is not None else torch.empty_like(w1)
        gb1 = sinks[1] if sinks[1] is not None else torch.empty_like(b1)
        gw2 = sinks[2] if sinks[2] is not None else torch.empty_like(w2)
        gb2 = sinks[3] if sinks[3] is not None else torch.empty(w2.shape[0], dtype=torch.float32, device=dev)
        nws = L.deno_fno_workspace_bytes(_C.byref(desc), _capi.FNO_WS_PROJ_BWD)
        ws = torch.empty(max(nws, 1), dtype=torch.uint8, device=dev)
        _poison(ws, gh, *[g for g, sk in zip((gw1, gb1, gw2, gb2), sinks) if sk is None])
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream(dev).cuda_stream
            _capi.check(L, L.deno_fno_proj_bwd(_C.byref(desc), gy.data_ptr(), h.data_ptr(), w1.data_ptr(),
                                               b1.data_ptr(), w2.data_ptr(), gh.data_ptr(), gw1.data_ptr(),
                                               gb1.data_ptr(), gw2.data_ptr(), gb2.data_ptr(), ws.data_ptr(),
                                               ws.numel(), st), "deno_fno_proj_bwd")
        _count_launches(L)
        return (gh,) + tuple(None if sk is not None else g for sk, g in zip(sinks, (gw1, gb1, gw2, gb2))) + (None, None)


def _fused_ends_enabled() -> bool:
    """``DENO_FUSED_ENDS=0`` keeps lift / projection on the torch-op path (read per call; A/B measurements)."""
    return _os.environ.get("DENO_FUSED_ENDS", "1") != "0"


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
        desc = self._ends_desc(x, grid)
        if desc is not None:
            h = _LiftFn.apply(x, grid, self.fc0.weight, self.fc0.bias, desc, torch.is_grad_enabled())
        elif x.is_cuda and _fused_ends_enabled():
            h = self._lift_channels_first(x, grid)
        else:
            h = _tall_linear(torch.cat((x, grid), dim=-1), self.fc0)
            h = h.movedim(-1, 1)
            if self.padding != 0:
                h = F.pad(h, [0, self.padding] * nd)
        # ``_stage_done(i)`` (set by training.TrainStep for data-parallel runs) is called from the backward pass as soon
        # as the gradient wrt the INPUT of stage i exists, i.e. stage i's own parameter gradients have been enqueued:
        # stage i < depth is convs[i], stage depth is the projection.  The lift finishes with backward() itself.
        notify = getattr(self, "_stage_done", None) if torch.is_grad_enabled() else None
        for i, layer in enumerate(self.convs):
            if notify is not None and h.requires_grad:
                h.register_hook(lambda g, i=i: notify(i))
            h = layer(h)
        if notify is not None and h.requires_grad:
            h.register_hook(lambda g, i=len(self.convs): notify(i))
        if desc is not None:
            return _ProjFn.apply(h, self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias, desc,
                                 torch.is_grad_enabled())
        if h.is_cuda and _fused_ends_enabled():
            return self._proj_channels_first(h)
        if self.padding != 0:
            h = h[(Ellipsis,) + (slice(None, -self.padding),) * nd]
        h = h.movedim(1, -1)
        return _tall_linear(F.gelu(_tall_linear(h, self.fc1)), self.fc2)

    # Widths outside the fused point-wise kernels (include/deno_fno.h: > 64, e.g. the width-128 Airfoil / Pipe stacks):
    # library GEMMs, but arranged channels-first so that no activation-sized permute / pad / crop copy is made around
    # them (the op-by-op order of FNOs.py:136-151 costs ~10 such passes per step at width 128).
    def _lift_channels_first(self, x, grid):
        """``pad(permute(fc0(cat(x, grid))))`` as ONE batched GEMM that writes the padded channels-first tensor: the few
        input features are permuted and zero-padded first (tiny), with a constant-one feature carrying the bias so that
        the padding stays exactly zero."""
        nd = self.ndim
        feats = torch.cat((x, grid, torch.ones_like(x[..., :1])), dim=-1).movedim(-1, 1)     # (B, K + 1, *spatial)
        if self.padding != 0:
            feats = F.pad(feats, [0, self.padding] * nd)
        w = torch.cat((self.fc0.weight, self.fc0.bias[:, None]), dim=1)                         # (C, K + 1)
        B = feats.shape[0]
        # bmm with the (stride-0) expanded weight: torch.matmul(2-D, 3-D) folds the batch into a transposed mm and pays an
        # activation-sized clone in forward and backward (tools/copy_hunt.py)
        h = torch.bmm(w.unsqueeze(0).expand(B, -1, -1), feats.reshape(B, feats.shape[1], -1))   # (B, C, S)
        return h.view((B, h.shape[1]) + tuple(feats.shape[2:]))

    def _proj_channels_first(self, h):
        """``fc2(gelu(fc1(permute(crop(h)))))`` on the padded channels-first tensor as it is (1x1 convs = batched GEMMs over
        the point axis); only the ``out_dim``-channel result is cropped and permuted."""
        nd = self.ndim
        B, C = h.shape[:2]
        psp = tuple(h.shape[2:])
        t = torch.baddbmm(self.fc1.bias.view(1, -1, 1), self.fc1.weight.unsqueeze(0).expand(B, -1, -1), h.reshape(B, C, -1))
        t = F.gelu(t)
        y = torch.baddbmm(self.fc2.bias.view(1, -1, 1), self.fc2.weight.unsqueeze(0).expand(B, -1, -1), t)
        y = y.view((B, y.shape[1]) + psp)
        if self.padding != 0:
            y = y[(Ellipsis,) + (slice(None, -self.padding),) * nd]
        return y.movedim(1, -1).contiguous()

    def _ends_desc(self, x, grid):
        """Descriptor of the fused lift / projection kernels, or None when this call stays on torch ops
        (CPU tensors fail later in the spectral layer; unusual dtypes, a grid that needs a gradient or a shape
        outside include/deno_fno.h's coverage take the op-by-op path)."""
        nd = self.ndim
        if not (_fused_ends_enabled() and x.is_cuda and grid.is_cuda and x.dtype == torch.float32
                and grid.dtype == torch.float32 and x.dim() == nd + 2 and grid.dim() == nd + 2
                and x.shape[:-1] == grid.shape[:-1] and not grid.requires_grad
                and self.fc0.weight.dtype == torch.float32
                and x.shape[-1] + grid.shape[-1] == self.fc0.in_features):
            return None
        desc = _capi.make_fno_desc(x.shape[0], self.width, self.hidden, x.shape[-1], grid.shape[-1], self.out_dim,
                                   tuple(x.shape[1:-1]), self.padding)
        return desc if _lib().deno_fno_supported(_C.byref(desc)) else None


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
