"""``spectral_conv`` - the autograd boundary between PyTorch and libdeno_spectral.so.

One call = one ``SpectralConv{1,2,3}d.forward`` of the reference
(/root/reference/Models/fno/spectral_layers.py:84-115, 183-208, 299-338); its backward is the
hand-written adjoint, not autograd through library ops.  PyTorch is used
here for device memory, streams and the autograd graph only.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _capi
from ._lib import lib

import os

_TWIDDLES: Dict[tuple, torch.Tensor] = {}
_LAUNCHES = {"count": 0}
_POISON = {"on": os.environ.get("DENO_POISON", "0") == "1"}
# DENO_BWD_OVERLAP=0: parameter-gradient kernels stay on the caller's stream (deno_spectral_bwd instead of _overlap)
_OVERLAP = {"on": os.environ.get("DENO_BWD_OVERLAP", "1") != "0"}
_OVERLAP_HANDLES: Dict[int, "_capi.Overlap"] = {}


def set_bwd_overlap(on: bool):
    """Run the parameter-gradient kernels of every layer's backward on a side stream next to the input-gradient
    chain (``deno_spectral_bwd_overlap``, include/deno_spectral.h).  On by default; ``DENO_BWD_OVERLAP=0`` turns it off."""
    _OVERLAP["on"] = bool(on)


def _overlap_handle(dev: torch.device):
    """Side stream + fork / join events of this device (created once by the library, owned here for the process's life)."""
    if not _OVERLAP["on"]:
        return None
    h = _OVERLAP_HANDLES.get(dev.index)
    if h is None:
        if torch.cuda.is_current_stream_capturing():
            return None                       # never create streams / events inside a capture; the warm-up steps do it
        L = lib()
        h = _capi.Overlap()
        _capi.check(L, L.deno_overlap_create(C.byref(h)), "deno_overlap_create")
        _OVERLAP_HANDLES[dev.index] = h
    return h


def set_poison(on: bool):
    """Debug switch (also ``DENO_POISON=1``): every buffer this module hands to the library as an OUTPUT or as
    scratch - workspace, ``y``, saved ``act'(z)`` and spectrum, ``gx``, parameter-gradient outputs that are not
    registered sinks - is filled with a NaN bit pattern before the call.  A kernel that reads something it (or an
    earlier kernel of the same call) did not write then produces NaNs instead of silently using whatever the caching
    allocator left there.  Costs one fill kernel per buffer; off by default."""
    _POISON["on"] = bool(on)


def _poison(*tensors):
    if not _POISON["on"]:
        return
    for t in tensors:
        if t is None:
            continue
        if t.dtype == torch.uint8:
            t.fill_(0xFF)                      # 0xFFFFFFFF is a quiet NaN as fp32
        elif t.is_complex():
            torch.view_as_real(t).fill_(float("nan"))
        else:
            t.fill_(float("nan"))


def reload_env():
    """Make the library read its ``DENO_*`` switches from the environment again (it reads them once)."""
    lib().deno_reload_env()


def launch_count() -> int:
    """Kernels enqueued through this module since import (bench.py's ``gpu_launches``)."""
    return _LAUNCHES["count"]


def _normalize_modes(modes, ndim: int) -> Tuple[int, ...]:
    if isinstance(modes, int):
        return (modes,) * ndim
    modes = tuple(int(m) for m in modes)
    if len(modes) != ndim:
        raise ValueError(f"expected {ndim} mode counts, got {modes}")
    return modes


def _desc_key(d: _capi.SpectralDesc):
    return (d.ndim, d.batch, d.cin, d.cout, tuple(d.spatial), tuple(d.modes))


def _fn(L, d, name: str):
    """Entry point ``name`` for this descriptor: the ``deno_spectral4_*`` family for the 4-D stand-alone layer."""
    if isinstance(d, _capi.SpectralDesc4):
        name = name.replace("deno_spectral_", "deno_spectral4_")
    return getattr(L, name)


def _twiddles(d, device: torch.device) -> torch.Tensor:
    """Constant DFT tables: filled on the host in fp64 by the library, uploaded once per shape."""
    key = (device.index, d.ndim, tuple(d.spatial), tuple(d.modes))
    t = _TWIDDLES.get(key)
    if t is None:
        L = lib()
        nbytes = _fn(L, d, "deno_spectral_twiddle_bytes")(C.byref(d))
        if nbytes == 0:
            raise _capi.DenoError(L.deno_last_error().decode())
        host = np.empty(nbytes // 4, dtype=np.float32)
        _capi.check(L, _fn(L, d, "deno_spectral_twiddle_fill")(C.byref(d), host.ctypes.data_as(C.c_void_p), nbytes),
                    "deno_spectral_twiddle_fill")
        t = torch.from_numpy(host).to(device)
        _TWIDDLES[key] = t
    return t


def _check_input(x: torch.Tensor, what: str):
    if not x.is_cuda:
        raise RuntimeError(f"{what} must be a CUDA tensor: deno4pytorch_b200 has no CPU path "
                           "(the reference's own torch.fft path covers CPU tensors)")
    if x.dtype != torch.float32:
        raise RuntimeError(f"{what} must be float32 (got {x.dtype}); the spectral path is fp32 in / fp32 out")


def _check_weight(w: torch.Tensor, cin: int, cout: int, modes: Tuple[int, ...], j: int):
    shape_c = (cin, cout) + modes
    if w.dtype == torch.complex64 and tuple(w.shape) == shape_c:
        return
    if w.dtype == torch.float32 and tuple(w.shape) == shape_c + (2,):
        return
    raise RuntimeError(f"weights{j + 1} must be complex64 {shape_c} or float32 {shape_c + (2,)}, "
                       f"got {w.dtype} {tuple(w.shape)}")


def grad_sink(param, like: Optional[torch.Tensor] = None):
    """Where the backward kernels may write this parameter's gradient DIRECTLY: the ``.grad`` view registered by
    ``parallel.FlatGradients.direct_writes()`` (one flat buffer for the optimiser and the all-reduce).  With a sink
    the backward returns ``None`` for the parameter, so autograd launches no ``grad += new`` kernel for it.  Only
    valid when the parameter is used once per backward pass (the registering side's responsibility)."""
    s = getattr(param, "_deno_grad_sink", None)
    if s is None or not s.is_contiguous():
        return None
    ref = param if like is None else like
    if s.dtype != ref.dtype or tuple(s.shape) != tuple(ref.shape) or s.device != ref.device:
        return None
    return s


def _claim_sink(param, like: Optional[torch.Tensor] = None):
    """``grad_sink`` for ONE pending backward: a direct write overwrites, so a parameter that is used a second time
    before the first use's backward has run (roll-out / BPTT, a loss that calls the model twice, micro-batch
    accumulation with retained graphs) must not take the sink again - that use returns its gradient to autograd,
    which accumulates it.  The first use's direct write then ZEROES nothing: to stay exact the first claim is
    revoked as well (``_deno_sink_shared``), i.e. once a parameter is seen in two live graphs every use accumulates
    through autograd until the sink is registered afresh (``FlatGradients.direct_writes``)."""
    if param is None:
        return None
    s = grad_sink(param, like)
    if s is None:
        return None
    if getattr(param, "_deno_sink_shared", False):
        return None
    if getattr(param, "_deno_sink_pending", 0) > 0:
        param._deno_sink_shared = True       # a second live use: the sink cannot be used for either of them
        return None
    param._deno_sink_pending = 1
    return s


def _resolve_sink(param, sink):
    """Backward-time view of a sink claimed in forward: None once the parameter has been seen in a second live graph
    (the other use accumulates through autograd into the same ``.grad`` memory, which a direct write would clobber)."""
    if sink is None or param is None or getattr(param, "_deno_sink_shared", False):
        return None
    return sink


def _release_sink(param):
    if param is not None and getattr(param, "_deno_sink_pending", 0) > 0:
        param._deno_sink_pending = 0


class _SpectralConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, lin_w, lin_b, modes, act, save, *weights):
        L = lib()
        _check_input(x, "input")
        ndim = x.dim() - 2
        if ndim not in (1, 2, 3, 4):
            raise RuntimeError("input must be (B, C, *spatial) with 1-4 spatial dims")
        B, cin = x.shape[:2]
        spatial = tuple(x.shape[2:])
        cout = lin_w.shape[0]
        if lin_w.shape[1] != cin or lin_w.numel() != cout * cin:
            raise RuntimeError(f"linear.weight {tuple(lin_w.shape)} does not match {cin} input channels")
        nblocks = 4 if ndim == 4 else 2 ** (ndim - 1)      # 4-D (FNO_HIT.py): both signs on the first two axes only
        if len(weights) != nblocks:
            raise RuntimeError(f"{ndim}-D layer needs {nblocks} weight blocks, got {len(weights)}")
        for j, w in enumerate(weights):
            _check_weight(w, cin, cout, modes, j)
        d = _capi.make_desc4(B, cin, cout, spatial, modes, act) if ndim == 4 else \
            _capi.make_desc(B, cin, cout, spatial, modes, act)

        # ``save`` = grad mode of the CALLER (inside Function.forward it is always off): under torch.no_grad()
        # (validation, roll-out inference) nothing is saved, so act'(z) and the spectrum are neither allocated nor written
        needs_grad = bool(save) and any(ctx.needs_input_grad)
        xc = x.contiguous()
        wc = [w.contiguous() for w in weights]
        lw = lin_w.contiguous()
        lb = None if lin_b is None else lin_b.contiguous()
        dev = x.device
        with torch.cuda.device(dev):
            tw = _twiddles(d, dev)
            y = torch.empty((B, cout) + spatial, dtype=torch.float32, device=dev)
            z = torch.empty_like(y) if needs_grad else None
            nmodes = _fn(L, d, "deno_spectral_num_modes")(C.byref(d))
            xhat = torch.empty((B, cin, nmodes), dtype=torch.complex64, device=dev) if needs_grad else None
            nws = _fn(L, d, "deno_spectral_workspace_bytes")(C.byref(d), 0)
            ws = torch.empty(max(nws, 1), dtype=torch.uint8, device=dev)
            _poison(ws, y, z, xhat)
            stream = torch.cuda.current_stream(dev).cuda_stream
            rc = _fn(L, d, "deno_spectral_fwd")(C.byref(d), xc.data_ptr(), _capi.pointer_array([w.data_ptr() for w in wc]),
                                     lw.data_ptr(), None if lb is None else lb.data_ptr(), tw.data_ptr(),
                                     y.data_ptr(), None if z is None else z.data_ptr(),
                                     None if xhat is None else xhat.data_ptr(), ws.data_ptr(), nws, stream)
            _capi.check(L, rc, "deno_spectral_fwd")
            _LAUNCHES["count"] += L.deno_spectral_last_launch_count()
        if needs_grad:
            ctx.save_for_backward(xc, z, xhat, lw, *wc)
            ctx.sinks = (_claim_sink(lin_w, lw), None if lin_b is None else _claim_sink(lin_b, lb),
                         [_claim_sink(w, c) for w, c in zip(weights, wc)])
            ctx.sink_owners = [lin_w, lin_b] + list(weights)
            ctx.desc = d
            ctx.has_bias = lin_b is not None
            ctx.weight_meta = [(w.dtype, tuple(w.shape)) for w in weights]
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        L = lib()
        xc, z, xhat, lw = ctx.saved_tensors[:4]
        wc = ctx.saved_tensors[4:]
        d = ctx.desc
        need_x, need_lw, need_lb = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        need_w = any(ctx.needs_input_grad[5:])
        dev = xc.device
        gy = gy.contiguous()
        with torch.cuda.device(dev):
            tw = _twiddles(d, dev)
            gx = torch.empty_like(xc) if need_x else None
            s_lw, s_lb, s_w = ctx.sinks
            owners = ctx.sink_owners
            s_lw, s_lb = _resolve_sink(owners[0], s_lw), _resolve_sink(owners[1], s_lb)
            s_w = [_resolve_sink(o, sk) for o, sk in zip(owners[2:], s_w)]
            gws = [sk if sk is not None else torch.empty(shape, dtype=dtype, device=dev)
                   for sk, (dtype, shape) in zip(s_w, ctx.weight_meta)] if need_w else None
            glw = (s_lw if s_lw is not None else torch.empty_like(lw)) if need_lw else None
            glb = (s_lb if s_lb is not None else torch.empty(lw.shape[0], dtype=torch.float32, device=dev)) \
                if (need_lb and ctx.has_bias) else None
            nws = _fn(L, d, "deno_spectral_workspace_bytes")(C.byref(d), 1)
            ws = torch.empty(max(nws, 1), dtype=torch.uint8, device=dev)
            _poison(ws, gx, glw if s_lw is None else None, glb if s_lb is None else None,
                    *([g for g, sk in zip(gws, s_w) if sk is None] if gws is not None else []))
            stream = torch.cuda.current_stream(dev).cuda_stream
            ov = _overlap_handle(dev)
            bwd = L.deno_spectral4_bwd if isinstance(d, _capi.SpectralDesc4) else L.deno_spectral_bwd_overlap
            rc = bwd(
                C.byref(d), gy.data_ptr(), xc.data_ptr(), z.data_ptr(), xhat.data_ptr(),
                _capi.pointer_array([w.data_ptr() for w in wc]), lw.data_ptr(), tw.data_ptr(),
                None if gx is None else gx.data_ptr(),
                None if gws is None else _capi.pointer_array([g.data_ptr() for g in gws]),
                None if glw is None else glw.data_ptr(), None if glb is None else glb.data_ptr(),
                ws.data_ptr(), nws, stream, None if ov is None else C.byref(ov))
            _capi.check(L, rc, "deno_spectral_bwd_overlap")
            _LAUNCHES["count"] += L.deno_spectral_last_launch_count()
        for owner in ctx.sink_owners:
            _release_sink(owner)
        # gradients written straight into a registered sink are not handed to autograd (nothing left to accumulate)
        grads_w = tuple(None if sk is not None else g for sk, g in zip(s_w, gws)) if gws is not None else (None,) * len(wc)
        return (gx, None if s_lw is not None else glw, None if s_lb is not None else glb, None, None, None) + grads_w


def spectral_conv(x: torch.Tensor, weights: Sequence[torch.Tensor], lin_w: torch.Tensor,
                  lin_b: Optional[torch.Tensor], modes, activation: str = "gelu") -> torch.Tensor:
    """``act(irfftn(corner_mix(rfftn(x))) + conv1x1(x) + bias)`` on the CUDA kernels.

    ``x`` (B, Cin, *spatial) fp32 CUDA; ``weights`` the reference's 1/2/4 blocks (complex64 or the
    fp32 ``(..., 2)`` real view; four spatial dims: the four blocks of ``SpectralConv4d``,
    /root/reference/Demo/Turbulence_3d+t/FNO_HIT.py:31-78); ``lin_w``/``lin_b`` the 1x1 conv parameters; ``activation`` one of
    identity/gelu/silu/relu/tanh/leakyrelu.  Differentiable w.r.t. x, weights, lin_w, lin_b.
    """
    modes = _normalize_modes(modes, x.dim() - 2)
    if activation not in _capi.ACTIVATION_IDS:
        raise ValueError(f"unknown activation {activation!r}")
    return _SpectralConvFn.apply(x, lin_w, lin_b, modes, activation, torch.is_grad_enabled(), *weights)


# ----------------------------------------------------------------------------------------------------------------
# Adaptive Fourier layers (include/deno_afno.h)
# ----------------------------------------------------------------------------------------------------------------
def afno_kept_modes(spatial: Sequence[int], hard_thresholding_fraction: float) -> int:
    """Bins of the half-spectrum axis the reference's ``[..., :kept_modes]`` slice leaves
    (/root/reference/Models/fno/spectral_layers.py:394-395, :470-471, :547-548): ``kept_modes`` is computed from the
    size of the WHOLE grid but applied to the last axis."""
    n = 1
    for v in spatial:
        n *= int(v)
    kept = int((n // 2 + 1) * hard_thresholding_fraction)
    return max(0, min(kept, int(spatial[-1]) // 2 + 1))


def _afno_twiddles(d: "_capi.AfnoDesc", device: torch.device) -> torch.Tensor:
    key = ("afno", device.index, d.ndim, d.channels, tuple(d.spatial), d.modes_last)
    t = _TWIDDLES.get(key)
    if t is None:
        L = lib()
        nbytes = L.deno_afno_twiddle_bytes(C.byref(d))
        if nbytes == 0:
            raise _capi.DenoError(L.deno_last_error().decode())
        host = np.zeros(nbytes // 4, dtype=np.float32)
        _capi.check(L, L.deno_afno_twiddle_fill(C.byref(d), host.ctypes.data_as(C.c_void_p), nbytes), "deno_afno_twiddle_fill")
        t = torch.from_numpy(host).to(device)
        _TWIDDLES[key] = t
    return t


def _check_afno_param(t: torch.Tensor, shape, name: str):
    if t.dtype != torch.complex64 or tuple(t.shape) != tuple(shape) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA complex64 tensor {tuple(shape)}, got {t.dtype} {tuple(t.shape)} on {t.device}")


class _AfnoFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, num_blocks, modes_last, act, lam, save):
        L = lib()
        _check_input(x, "input")
        ndim = x.dim() - 2
        if ndim not in (1, 2, 3):
            raise RuntimeError("input must be (B, C, *spatial) with 1-3 spatial dims")
        B, Cn = x.shape[:2]
        if Cn % num_blocks != 0:
            raise RuntimeError(f"hidden_size {Cn} should be divisble by num_blocks {num_blocks}")
        bs = Cn // num_blocks
        _check_afno_param(w1, (num_blocks, bs, bs), "weights1")
        _check_afno_param(b1, (num_blocks, bs), "bias1")
        _check_afno_param(w2, (num_blocks, bs, bs), "weights2")
        _check_afno_param(b2, (num_blocks, bs), "bias2")
        d = _capi.make_afno_desc(B, Cn, num_blocks, tuple(x.shape[2:]), modes_last, act, lam)
        needs_grad = bool(save) and any(ctx.needs_input_grad)
        dev = x.device
        with torch.cuda.device(dev):
            xc = x.contiguous()
            p = [t.detach().contiguous() for t in (w1, b1, w2, b2)]
            tw = _afno_twiddles(d, dev)
            y = torch.empty_like(xc)
            nmodes = L.deno_afno_num_modes(C.byref(d))
            if nmodes == 0:
                raise _capi.DenoError(L.deno_last_error().decode())
            xhat = torch.empty((B, Cn, nmodes), dtype=torch.complex64, device=dev) if needs_grad else None
            nws = L.deno_afno_workspace_bytes(C.byref(d), 0)
            if nws == 0:
                raise _capi.DenoError(L.deno_last_error().decode())
            ws = torch.empty(nws, dtype=torch.uint8, device=dev)
            _poison(ws, y, xhat)
            stream = torch.cuda.current_stream(dev).cuda_stream
            rc = L.deno_afno_fwd(C.byref(d), xc.data_ptr(), p[0].data_ptr(), p[1].data_ptr(), p[2].data_ptr(),
                                 p[3].data_ptr(), tw.data_ptr(), y.data_ptr(), None if xhat is None else xhat.data_ptr(),
                                 ws.data_ptr(), nws, stream)
            _capi.check(L, rc, "deno_afno_fwd")
            _LAUNCHES["count"] += L.deno_spectral_last_launch_count()
        if needs_grad:
            ctx.save_for_backward(xhat, *p)
            ctx.desc = d
        return y

    @staticmethod
    def backward(ctx, gy):
        L = lib()
        xhat, w1, b1, w2, b2 = ctx.saved_tensors
        d = ctx.desc
        dev = gy.device
        need_x = ctx.needs_input_grad[0]
        need_p = any(ctx.needs_input_grad[1:5])
        with torch.cuda.device(dev):
            gy = gy.contiguous()
            gx = torch.empty_like(gy) if need_x else None
            gp = [torch.empty_like(t) for t in (w1, b1, w2, b2)] if need_p else [None] * 4
            nws = L.deno_afno_workspace_bytes(C.byref(d), 1)
            ws = torch.empty(max(nws, 1), dtype=torch.uint8, device=dev)
            _poison(ws, gx, *gp)
            stream = torch.cuda.current_stream(dev).cuda_stream
            rc = L.deno_afno_bwd(C.byref(d), gy.data_ptr(), xhat.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(),
                                 b2.data_ptr(), _afno_twiddles(d, dev).data_ptr(), None if gx is None else gx.data_ptr(),
                                 *[None if g is None else g.data_ptr() for g in gp], ws.data_ptr(), nws, stream)
            _capi.check(L, rc, "deno_afno_bwd")
            _LAUNCHES["count"] += L.deno_spectral_last_launch_count()
        return (gx, gp[0], gp[1], gp[2], gp[3], None, None, None, None, None)


def afno_filter(x: torch.Tensor, weights1: torch.Tensor, bias1: torch.Tensor, weights2: torch.Tensor, bias2: torch.Tensor,
                *, num_blocks: int, sparsity_threshold: float = 0.01, hard_thresholding_fraction: float = 1,
                activation: str = "gelu") -> torch.Tensor:
    """``x + irfftn(softshrink(W2^T act(W1^T rfftn(x) + b1) + b2))`` (ortho transforms, block-diagonal complex weights)
    on the CUDA kernels: one ``AdaptiveFourier{1,2,3}d.forward`` of the reference
    (/root/reference/Models/fno/spectral_layers.py:383-409, :457-487, :533-565).  ``x`` (B, C, *spatial) fp32 CUDA;
    weights ``(num_blocks, C/num_blocks, C/num_blocks)`` and biases ``(num_blocks, C/num_blocks)`` complex64.
    Differentiable w.r.t. x and the four parameters."""
    if activation not in _capi.ACTIVATION_IDS:
        raise ValueError(f"unknown activation {activation!r}")
    _check_input(x, "input")
    kept = afno_kept_modes(tuple(x.shape[2:]), hard_thresholding_fraction)
    if kept == 0:
        # every bin is masked out: the filter branch is soft-shrink(0) = 0.  The reference still runs its einsums on
        # empty slices, so the four parameters receive ZERO gradients there (not None): keep them in the graph
        zero = sum(torch.view_as_real(p)[..., :0].sum() for p in (weights1, bias1, weights2, bias2))
        return x + zero
    return _AfnoFn.apply(x, weights1, bias1, weights2, bias2, int(num_blocks), kept, activation, float(sparsity_threshold),
                         torch.is_grad_enabled())
