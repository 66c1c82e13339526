"""Batch-sharded data parallelism for the FNO stacks: one process per GPU, NCCL all-reduce of one
flat fp32 gradient buffer.

Replaces, for this path, the reference's ``setup_DDP`` + ``DistributedDataParallel`` wrapper
(/root/reference/Utilizes/parallel_run.py:18-45, /root/reference/Demo/Turbulence_3d+t/run_train_DDP.py:272).
Differences, all deliberate:
  * parameters' ``.grad`` are views into ONE flat fp32 buffer (complex64 spectral weights as their
    real view, which is also what c10d reduces), so the whole model needs a single collective and
    no bucket copies;
  * no per-step ``barrier()`` and no per-step scalar all-reduce + ``.item()`` (run_train_DDP.py:87-93):
    the loss stays on the device.
The data path has no other collective: samples are independent through the whole model.
"""
from __future__ import annotations

import os
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


def setup_ddp(backend: Optional[str] = None, verbose: bool = False):
    """Env-var bootstrap with the reference's return convention ``(rank, local_rank, world_size, device)``
    (parallel_run.py:18-45).  Works under ``torchrun`` / ``torch.distributed.run``; without the
    launcher's variables it returns the single-process values and does not initialise a group."""
    if "RANK" not in os.environ or "WORLD_SIZE" not in os.environ:
        device = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
        return 0, 0, 1, device
    rank = int(os.environ["RANK"])
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world_size = int(os.environ["WORLD_SIZE"])
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    use_cuda = torch.cuda.is_available()
    if backend is None:
        backend = "nccl" if use_cuda else "gloo"
    if use_cuda:
        torch.cuda.set_device(local_rank)
        device = torch.device("cuda", local_rank)
    else:
        device = torch.device("cpu")
    if not dist.is_initialized():
        kwargs = {}
        if use_cuda and backend == "nccl":
            kwargs["device_id"] = device
        dist.init_process_group(backend=backend, rank=rank, world_size=world_size, **kwargs)
    if verbose:
        print(f"rank {rank}/{world_size} local_rank {local_rank} device {device} backend {backend}", flush=True)
    return rank, local_rank, world_size, device


class FlatGradients:
    """Owns one flat fp32 buffer and makes every parameter's ``.grad`` a view into it."""

    def __init__(self, params: Iterable[torch.nn.Parameter], alloc=None):
        """``alloc(numel, device) -> flat fp32 tensor`` overrides where the flat buffer lives (symmetric memory for the
        fused multicast exchange); default is an ordinary device tensor."""
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        device = self.params[0].device
        sizes = []
        for p in self.params:
            if p.dtype not in (torch.float32, torch.complex64):
                raise TypeError(f"unsupported parameter dtype {p.dtype}")
            n = p.numel() * (2 if p.is_complex() else 1)
            sizes.append((n + 3) // 4 * 4)  # keep every view 16-byte aligned
        if alloc is not None:
            self.flat = alloc(sum(sizes), device)
            self.flat.zero_()
        else:
            self.flat = torch.zeros(sum(sizes), dtype=torch.float32, device=device)
        self.spans = {}            # id(param) -> (offset, padded length) inside ``flat``
        off = 0
        for p, n in zip(self.params, sizes):
            self.spans[id(p)] = (off, n)
            real_n = p.numel() * (2 if p.is_complex() else 1)
            seg = self.flat[off:off + real_n]
            if p.is_complex():
                p.grad = torch.view_as_complex(seg.view(-1, 2)).view(p.shape)
            else:
                p.grad = seg.view(p.shape)
            off += n

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * 4

    def direct_writes(self, enable: bool = True):
        """Let the backward kernels write parameter gradients straight into this buffer (``functional.grad_sink``)
        instead of returning them to autograd, which would launch one ``grad += new`` kernel per parameter.
        Valid only while every parameter is used ONCE per backward pass (a plain model call; not a roll-out that
        calls the model several times before ``backward``): a direct write overwrites, it does not accumulate."""
        for p in self.params:
            p._deno_grad_sink = p.grad if enable else None
        self.reset_pending()

    def reset_pending(self):
        """Forget claims left behind by a forward whose backward never ran (an evaluation with grad mode on) and the
        "used in two live graphs" marks; called when sinks are (de)registered and at the start of every eager step."""
        for p in self.params:
            p._deno_sink_pending = 0
            p._deno_sink_shared = False

    def zero(self):
        self.flat.zero_()

    def span_of(self, params: Iterable[torch.nn.Parameter]) -> torch.Tensor:
        """The contiguous slice of the flat buffer covering ``params`` (which must be adjacent in it)."""
        spans = sorted(self.spans[id(p)] for p in params if id(p) in self.spans)
        if not spans:
            return self.flat[:0]
        for (o0, n0), (o1, _) in zip(spans, spans[1:]):
            if o0 + n0 != o1:
                raise ValueError("parameters are not adjacent in the flat gradient buffer")
        return self.flat[spans[0][0]:spans[-1][0] + spans[-1][1]]

    def all_reduce_mean(self, group=None, async_op: bool = False):
        """Sum over ranks then divide by world size (DDP's gradient averaging)."""
        if not dist.is_initialized() or dist.get_world_size(group) == 1:
            return None
        world = dist.get_world_size(group)
        self.flat.div_(world)
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)


class FlatParameters:
    """Moves every trainable parameter into ONE flat fp32 buffer (``p.data`` becomes a view; names, shapes and
    dtypes - hence the state dict - are unchanged) and pairs it with a ``FlatGradients`` buffer of the same
    layout, so the optimiser can be a single fused kernel over flat arrays."""

    def __init__(self, params: Iterable[torch.nn.Parameter], alloc=None):
        self.grads = FlatGradients(params, alloc=alloc)
        self.params = self.grads.params
        if alloc is not None:
            self.flat = alloc(self.grads.flat.numel(), self.grads.flat.device)
            self.flat.zero_()
        else:
            self.flat = torch.zeros_like(self.grads.flat)
        off = 0
        with torch.no_grad():
            for p in self.params:
                real_n = p.numel() * (2 if p.is_complex() else 1)
                seg = self.flat[off:off + real_n]
                src = torch.view_as_real(p.data).reshape(-1) if p.is_complex() else p.data.reshape(-1)
                seg.copy_(src)
                p.data = torch.view_as_complex(seg.view(-1, 2)).view(p.shape) if p.is_complex() else seg.view(p.shape)
                off += (real_n + 3) // 4 * 4


class FlatAdam:
    """torch.optim.Adam's update (no amsgrad) as ONE kernel launch over ``FlatParameters`` (libdeno_spectral's
    ``deno_adam_step``).  The step counter lives on the device, so ``step()`` can be captured in a CUDA graph."""

    def __init__(self, flat: FlatParameters, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.flat = flat
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.exp_avg = torch.zeros_like(flat.flat)
        self.exp_avg_sq = torch.zeros_like(flat.flat)
        self.step_t = torch.zeros(1, dtype=torch.float32, device=flat.flat.device)

    def zero_grad(self):
        self.flat.grads.zero()

    def step(self):
        from . import _capi
        from ._lib import lib
        L = lib()
        f = self.flat
        self.step_t.add_(1.0)
        with torch.cuda.device(f.flat.device):
            rc = L.deno_adam_step(f.flat.data_ptr(), f.grads.flat.data_ptr(), self.exp_avg.data_ptr(),
                                  self.exp_avg_sq.data_ptr(), f.flat.numel(), self.lr, self.betas[0], self.betas[1],
                                  self.eps, self.weight_decay, self.step_t.data_ptr(),
                                  torch.cuda.current_stream(f.flat.device).cuda_stream)
        _capi.check(L, rc, "deno_adam_step")
        from .functional import _LAUNCHES
        _LAUNCHES["count"] += 1


class SymmetricFlatBuffers:
    """Allocator + rendezvous of the flat buffers in SYMMETRIC memory (same size and offset on every rank, one
    multicast mapping each) for the fused all-reduce + Adam kernel (include/deno_dp.h).  PyTorch's symmetric-memory
    module is the plumbing: allocation, handle exchange, multicast binding, the signal pads."""

    def __init__(self, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self._symm_mem = symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.tensors: List[torch.Tensor] = []
        self.handles = []

    def alloc(self, numel: int, device) -> torch.Tensor:
        t = self._symm_mem.empty(numel, dtype=torch.float32, device=device)
        self.tensors.append(t)
        return t

    def rendezvous(self):
        """Collective.  Returns False (on every rank, after agreeing) when any rank lacks multicast support."""
        ok = 1
        try:
            self.handles = [self._symm_mem.rendezvous(t, self.group) for t in self.tensors]
            for h in self.handles:
                if not h.has_multicast_support or int(h.multicast_ptr) == 0:
                    ok = 0
        except Exception:
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=self.tensors[0].device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        return bool(flag.item())


class FusedDPAdam:
    """``FlatAdam`` for data-parallel training with the collective INSIDE the optimiser kernel: all-reduce(mean) of the
    flat gradient through NVLink multicast, Adam on this rank's slice, multicast store of the new parameters
    (``deno_dp_fused_adam``, csrc/dp_fused.cuh).  Replaces ``ncclAllReduce`` + ``adam_flat``.  The Adam moments are
    sharded: ``exp_avg`` / ``exp_avg_sq`` hold valid values for this rank's slice only."""

    def __init__(self, flat: FlatParameters, buffers: SymmetricFlatBuffers, lr=1e-3, betas=(0.9, 0.999), eps=1e-8,
                 weight_decay=0.0, write_back_grads=False):
        self.flat = flat
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.exp_avg = torch.zeros_like(flat.flat)
        self.exp_avg_sq = torch.zeros_like(flat.flat)
        self.step_t = torch.zeros(1, dtype=torch.float32, device=flat.flat.device)
        g_hdl, p_hdl = buffers.handles[0], buffers.handles[1]
        assert buffers.tensors[0].data_ptr() == flat.grads.flat.data_ptr() and buffers.tensors[1].data_ptr() == flat.flat.data_ptr()
        self._g_hdl, self._p_hdl = g_hdl, p_hdl          # keep the mappings alive
        self.rank, self.world = g_hdl.rank, g_hdl.world_size
        self.g_mc = int(g_hdl.multicast_ptr) + int(getattr(g_hdl, "offset", 0) or 0)
        self.p_mc = int(p_hdl.multicast_ptr) + int(getattr(p_hdl, "offset", 0) or 0)
        self.signal_pads = int(g_hdl.signal_pad_ptrs_dev)
        if g_hdl.signal_pad_size < 2048:             # 512 slots: up to 64 blocks x 8 ranks (include/deno_dp.h)
            raise RuntimeError("signal pad too small for the fused exchange")
        self.write_back_grads = bool(write_back_grads)

    def zero_grad(self):
        self.flat.grads.zero()

    def step(self):
        from . import _capi
        from ._lib import lib
        L = lib()
        f = self.flat
        self.step_t.add_(1.0)
        with torch.cuda.device(f.flat.device):
            rc = L.deno_dp_fused_adam(f.flat.data_ptr(), self.p_mc, self.g_mc, self.exp_avg.data_ptr(),
                                      self.exp_avg_sq.data_ptr(), f.flat.numel(), self.rank, self.world,
                                      self.signal_pads, int(self.write_back_grads), self.lr, self.betas[0],
                                      self.betas[1], self.eps, self.weight_decay, self.step_t.data_ptr(),
                                      torch.cuda.current_stream(f.flat.device).cuda_stream)
        _capi.check(L, rc, "deno_dp_fused_adam")
        from .functional import _LAUNCHES
        _LAUNCHES["count"] += 1


def shard_batch(t: torch.Tensor, rank: int, world_size: int) -> torch.Tensor:
    """This rank's contiguous slice of a global batch (batch-sharded DP; no data-path collective)."""
    n = t.shape[0]
    if n % world_size != 0:
        raise ValueError(f"global batch {n} not divisible by world size {world_size}")
    per = n // world_size
    return t[rank * per:(rank + 1) * per]
