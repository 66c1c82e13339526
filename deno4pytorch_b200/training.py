"""``TrainStep`` - one optimiser step of an FNO stack, the way every reference demo trains
(forward, MSE, backward, Adam: /root/reference/Demo/Darcy_2d/run_FNO&UNet.py:43-69,206-209),
captured once into a CUDA graph and replayed.

The reference loop is launch-bound on small grids (each step issues a few hundred tiny library
kernels from Python); the captured graph removes the per-launch host cost.  Under data
parallelism the gradient all-reduce (deno4pytorch_b200.parallel.FlatGradients) runs between the
backward graph and the optimiser graph on the same stream.
"""
from __future__ import annotations

import os
from typing import Callable, Optional

import torch
import torch.distributed as dist
import torch.nn.functional as F

from .parallel import FlatAdam, FlatGradients, FlatParameters, FusedDPAdam, SymmetricFlatBuffers


def _all_gradients_direct(model, x, grid) -> bool:
    """True when every parameter gradient of ``model`` is written by this library's backward kernels themselves
    (``functional.grad_sink``), i.e. is complete the moment a stage's backward function returns.  Gradients that
    still go through autograd's AccumulateGrad nodes (op-by-op ends, the dropout / return_freq branches of the layer)
    have no such ordering, and the per-stage all-reduce must not be used."""
    if not (hasattr(model, "convs") and hasattr(model, "_ends_desc")):
        return False
    if model._ends_desc(x, grid) is None:
        return False
    for layer in model.convs:
        if getattr(layer, "return_freq", False):
            return False
        drop = getattr(layer, "dropout", None)
        if getattr(layer, "_dropout_in_forward", False) and drop is not None and getattr(drop, "p", 0.0) > 0:
            return False
        if not callable(getattr(layer, "_weights", None)):
            return False
    return True


class TrainStep:
    def __init__(self, model: torch.nn.Module, x: torch.Tensor, grid: torch.Tensor, target: torch.Tensor, *,
                 lr: float = 1e-3, betas=(0.7, 0.9), weight_decay: float = 1e-4,
                 loss_fn: Callable = F.mse_loss, forward_fn: Optional[Callable] = None,
                 use_graph: bool = True, warmup: int = 3, group=None, fused_adam: bool = True,
                 overlap_allreduce: Optional[bool] = None):
        """``x``/``grid``/``target`` are example device tensors fixing the shapes; later calls copy new
        data into static buffers.  ``forward_fn(model, x, grid, target) -> loss`` overrides the default
        single model call (e.g. the 10-step autoregressive rollout of Demo/Turbulence_2d+t).

        Construction runs ``warmup`` throw-away iterations on the example batch (lazy initialisation, then graph
        capture) and then RESTORES the parameters and zeroes the optimiser state: the model leaves the constructor
        with the weights it came in with, and the first ``step()`` is Adam's step 1."""
        self.model = model
        self.loss_fn = loss_fn
        self.forward_fn = forward_fn
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.sx, self.sg, self.st = x.clone(), grid.clone(), target.clone()
        self._fused_dp = False
        self.flat_params = None
        if fused_adam and x.is_cuda:
            # Data parallel on an NVSwitch box: gradients and parameters live in symmetric memory and ONE kernel does
            # all-reduce + Adam + parameter broadcast over NVLink multicast (DENO_DP_FUSED=0: NCCL all-reduce + adam_flat)
            want_fused = (self.world > 1 and dist.get_backend(group) == "nccl" and not overlap_allreduce
                          and os.environ.get("DENO_DP_FUSED", "1") == "1" and os.environ.get("DENO_DP_OVERLAP", "0") != "1")
            buffers = None
            if want_fused:
                try:
                    buffers = SymmetricFlatBuffers(group)
                    self.flat_params = FlatParameters(model.parameters(), alloc=buffers.alloc)
                except Exception:
                    buffers = None
                flag = torch.tensor([1 if buffers is not None else 0], dtype=torch.int32, device=x.device)
                dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)      # every rank or none
                if flag.item() == 0:
                    if buffers is not None:                                    # move the parameters back to plain memory
                        self.flat_params = FlatParameters(model.parameters())
                    buffers = None
                elif not buffers.rendezvous():
                    self.flat_params = FlatParameters(model.parameters())
                    buffers = None
            if self.flat_params is None:
                self.flat_params = FlatParameters(model.parameters())
            self.grads = self.flat_params.grads
            if buffers is not None:
                self.opt = FusedDPAdam(self.flat_params, buffers, lr=lr, betas=betas, weight_decay=weight_decay,
                                       write_back_grads=os.environ.get("DENO_DP_FUSED_WRITE_GRADS", "0") == "1")
                self._fused_dp = True
            else:
                self.opt = FlatAdam(self.flat_params, lr=lr, betas=betas, weight_decay=weight_decay)
        else:
            self.grads = FlatGradients(model.parameters())
            self.opt = torch.optim.Adam(model.parameters(), lr=lr, betas=betas, weight_decay=weight_decay,
                                        capturable=use_graph, foreach=True)
        # one model call per step: parameter gradients go straight into the flat buffer (no per-parameter add kernels)
        self.grads.direct_writes(forward_fn is None and x.is_cuda)
        # ... and when EVERY gradient is written that way the flat buffer is fully overwritten each step: no memset
        self._skip_zero = bool(forward_fn is None and x.is_cuda and _all_gradients_direct(model, x, grid))
        # Data parallel on GPUs with the stock single-call step: the gradient all-reduce is issued per stage from inside
        # the backward pass on a side stream (projection first, then the spectral layers last to first, the lift after
        # backward() returns), so it overlaps the rest of the backward, and the whole step is ONE captured graph.
        if overlap_allreduce is None:     # opt-in (DENO_DP_OVERLAP=1) until it has been proven under graph capture on NCCL
            overlap_allreduce = os.environ.get("DENO_DP_OVERLAP", "0") == "1"
        self._overlap = bool(overlap_allreduce and not self._fused_dp and self.world > 1 and x.is_cuda and forward_fn is None
                             and dist.get_backend(group) == "nccl" and _all_gradients_direct(model, x, grid))
        self._ar_stream = torch.cuda.Stream() if self._overlap else None
        # NCCL without the per-stage overlap: ONE averaged all-reduce of the flat buffer, captured in the same graph as
        # the backward and the optimiser (DENO_DP_INGRAPH=0 keeps it between two graphs, as gloo / eager runs do)
        self._ingraph = bool(not self._overlap and not self._fused_dp and self.world > 1 and x.is_cuda and use_graph
                             and dist.get_backend(group) == "nccl" and os.environ.get("DENO_DP_INGRAPH", "1") == "1")
        if self._overlap:
            stages = [list(layer.parameters()) for layer in model.convs]
            stages.append(list(model.fc1.parameters()) + list(model.fc2.parameters()))
            self._stage_grads = [self.grads.span_of(ps) for ps in stages]
            self._lift_grads = self.grads.span_of(list(model.fc0.parameters()))
            covered = sum(t.numel() for t in self._stage_grads) + self._lift_grads.numel()
            if covered != self.grads.flat.numel():
                raise RuntimeError("stage buckets do not cover the flat gradient buffer")
            model._stage_done = self._reduce_stage
        self.loss = torch.zeros((), device=x.device)
        self._copy_stream = None          # input pipeline (prefetch): created on first use
        self._staged = None
        self._sets = None                 # two static input sets + two graphs once prefetch() is used under graph replay
        self._graphs = None
        self._active = 0
        self._loss_slots = None           # pipelined loss read-back (enable_loss_readback)
        self._nsteps = 0
        self.use_graph = use_graph
        self._g_bwd = self._g_opt = None
        if use_graph:
            self._capture(warmup)

    # ------------------------------------------------------------------------------------
    def _reduce_bucket(self, t: torch.Tensor):
        """Average ``t`` over the ranks on the side stream, ordered after everything enqueued so far on this one."""
        if t.numel() == 0:
            return
        self._ar_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self._ar_stream):
            dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group)

    def _reduce_stage(self, i: int):
        self._reduce_bucket(self._stage_grads[i])

    def _forward_backward(self):
        self.grads.reset_pending()       # claims left by a forward whose backward never ran (functional._claim_sink)
        if not self._skip_zero:
            self.grads.zero()
        if self.forward_fn is not None:
            loss = self.forward_fn(self.model, self.sx, self.sg, self.st)
        else:
            loss = self.loss_fn(self.model(self.sx, self.sg), self.st)
        loss.backward()
        if self._overlap:
            self._reduce_bucket(self._lift_grads)
            torch.cuda.current_stream().wait_stream(self._ar_stream)
        elif self._ingraph:
            dist.all_reduce(self.grads.flat, op=dist.ReduceOp.AVG, group=self.group)
        self.loss.copy_(loss.detach())

    def _snapshot(self):
        """Parameters before the warm-up iterations (the optimiser TrainStep creates starts from zero moments)."""
        with torch.no_grad():
            if self.flat_params is not None:
                return self.flat_params.flat.clone()
            return [p.detach().clone() for p in self.model.parameters()]

    def _restore(self, snap):
        """Undo the warm-up: parameters back to the snapshot, Adam moments and step count back to zero - IN PLACE, the
        captured graph keeps pointing at these tensors.  After construction the model and the optimiser are exactly
        as the caller handed them over; the first ``step()`` is the first update (bias-correction step 1)."""
        with torch.no_grad():
            if self.flat_params is not None:
                self.flat_params.flat.copy_(snap)
                self.opt.exp_avg.zero_()
                self.opt.exp_avg_sq.zero_()
                self.opt.step_t.zero_()
            else:
                for p, q in zip(self.model.parameters(), snap):
                    p.copy_(q)
                for st in self.opt.state.values():
                    for v in st.values():
                        if torch.is_tensor(v):
                            v.zero_()
            self.grads.zero()
            self.loss.zero_()

    def _capture(self, warmup: int):
        snap = self._snapshot()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):   # also fills the twiddle cache and cuBLAS workspaces
                self._forward_backward()
                if not (self._overlap or self._ingraph or self._fused_dp):
                    self.grads.all_reduce_mean(self.group)
                self.opt.step()
            self._restore(snap)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._g_bwd = torch.cuda.CUDAGraph()
        if self.world == 1 or self._overlap or self._ingraph or self._fused_dp:
            # no separate collective between backward and optimiser: the whole step is one graph launch
            with torch.cuda.graph(self._g_bwd):
                self._forward_backward()
                self.opt.step()
            return
        with torch.cuda.graph(self._g_bwd):
            self._forward_backward()
        self._g_opt = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._g_opt):
            self.opt.step()

    def allreduce_description(self) -> str:
        """How this step combines gradients across ranks (bench.py's ``config.allreduce``)."""
        if self.world == 1:
            return "none"
        if self._fused_dp:
            return ("fused all-reduce + Adam kernel over NVLink multicast (multimem.ld_reduce of the flat gradient, "
                    "update, multimem.st of the new parameters) inside the step graph")
        if self._overlap:
            return "per stage, overlapped with the backward, inside the graph"
        if self._ingraph:
            return "one flat averaged NCCL all-reduce inside the step graph"
        return "one flat NCCL all-reduce between the backward and the optimiser graph"

    def close(self):
        """Drop the captured graphs.  Needed before ``dist.destroy_process_group()`` when the all-reduce was captured
        (a live graph keeps NCCL kernels referenced and communicator teardown waits for them forever)."""
        torch.cuda.synchronize()
        self._g_bwd = self._g_opt = None
        self._graphs = None
        self._sets = None
        self.use_graph = False
        # the parameters keep their ``.grad`` views but no longer advertise them as direct-write targets: any other
        # backward on this model (gradient accumulation, a second loss) goes through autograd's accumulation again
        self.grads.direct_writes(False)

    def __del__(self):
        try:
            self.grads.direct_writes(False)
        except Exception:      # interpreter shutdown
            pass

    # ------------------------------------------------------------------------------------
    def load(self, x=None, grid=None, target=None, non_blocking: bool = True):
        """Copy a new batch (host or device tensors) into the static buffers."""
        if x is not None:
            self.sx.copy_(x, non_blocking=non_blocking)
        if grid is not None:
            self.sg.copy_(grid, non_blocking=non_blocking)
        if target is not None:
            self.st.copy_(target, non_blocking=non_blocking)

    def prefetch(self, x, grid, target):
        """Input pipeline: start the host-to-device copy of the NEXT batch (pinned host tensors) on a copy stream so it
        overlaps the step that is running; the next ``step()`` without arguments runs on it.

        Under graph replay the batch lands DIRECTLY in static inputs: the step is captured a second time on a second
        set of input buffers (same memory pool, the two graphs never run concurrently), ``prefetch`` fills the set the
        running step is not using and ``step()`` replays the graph that reads it - no staging buffers, no
        device-to-device copies.  Eager steps (``use_graph=False``) stage and copy."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream()
            self._staged_ready = torch.cuda.Event()
            if self.use_graph and self._g_opt is None and self._g_bwd is not None:
                self._capture_second_set()
            if self._sets is None:
                self._stage_bufs = (torch.empty_like(self.sx), torch.empty_like(self.sg), torch.empty_like(self.st))
                self._stage_free = torch.cuda.Event()
                self._stage_free.record()
        if self._sets is not None:
            tgt = 1 - self._active                              # the set the running / last step does not read
            self._copy_stream.wait_event(self._set_free[tgt])   # its previous step has finished with it
            with torch.cuda.stream(self._copy_stream):
                for dst, src in zip(self._sets[tgt], (x, grid, target)):
                    dst.copy_(src, non_blocking=True)
                self._staged_ready.record()
            self._staged = tgt + 1                              # truthy: set index + 1
            return
        self._copy_stream.wait_event(self._stage_free)      # the previous staged batch has been consumed
        with torch.cuda.stream(self._copy_stream):
            for dst, src in zip(self._stage_bufs, (x, grid, target)):
                dst.copy_(src, non_blocking=True)
            self._staged_ready.record()
        self._staged = True

    def _capture_second_set(self):
        """Second capture of the whole step on a second set of static inputs (shares the first graph's memory pool)."""
        first = (self.sx, self.sg, self.st)
        second = tuple(torch.empty_like(t) for t in first)
        for d, s_ in zip(second, first):
            d.copy_(s_)
        torch.cuda.synchronize()
        g2 = torch.cuda.CUDAGraph()
        self.sx, self.sg, self.st = second
        try:
            with torch.cuda.graph(g2, pool=self._g_bwd.pool()):
                self._forward_backward()
                self.opt.step()
        finally:
            self.sx, self.sg, self.st = first
        self._sets = [first, second]
        self._graphs = [self._g_bwd, g2]
        self._set_free = [torch.cuda.Event(), torch.cuda.Event()]
        for ev in self._set_free:
            ev.record()
        self._active = 0

    def enable_loss_readback(self):
        """Every step additionally copies its loss to pinned host memory (asynchronously, two slots);
        ``read_loss(lag)`` then returns the loss of the step ``lag`` calls back after waiting for THAT copy only, so a
        training loop can log step k - 1 while step k is already running (no GPU idle time between steps)."""
        if self._loss_slots is None:
            self._loss_slots = [(torch.zeros((), dtype=torch.float32).pin_memory(), torch.cuda.Event()) for _ in range(2)]

    def read_loss(self, lag: int = 0) -> float:
        """Host value of the loss of the step ``lag`` calls back (0: the latest; 1: the one before).  Needs
        ``enable_loss_readback()``; waits only for that step's device-to-host copy."""
        if self._loss_slots is None or lag not in (0, 1) or self._nsteps - 1 - lag < 0:
            raise RuntimeError("read_loss: enable_loss_readback() first; lag must be 0 or 1 with enough steps run")
        buf, ev = self._loss_slots[(self._nsteps - 1 - lag) % 2]
        ev.synchronize()
        return float(buf)

    def _consume_staged(self):
        cur = torch.cuda.current_stream()
        cur.wait_event(self._staged_ready)
        self.sx.copy_(self._stage_bufs[0])
        self.sg.copy_(self._stage_bufs[1])
        self.st.copy_(self._stage_bufs[2])
        self._stage_free.record(cur)
        self._staged = None

    def __call__(self, x=None, grid=None, target=None) -> torch.Tensor:
        """Runs one optimiser step; returns the (device) loss of this rank's shard.  Without arguments it runs on
        the batch staged by ``prefetch`` (if any), else on the batch already in the static buffers."""
        if x is None and grid is None and target is None and self._staged and self._sets is not None:
            # the staged batch already sits in the other graph's static inputs: switch graphs
            self._active = self._staged - 1
            self._staged = None
            cur = torch.cuda.current_stream()
            cur.wait_event(self._staged_ready)
            self.sx, self.sg, self.st = self._sets[self._active]
            self._graphs[self._active].replay()
            self._set_free[self._active].record(cur)
        else:
            if x is None and grid is None and target is None and self._staged:
                self._consume_staged()
            else:
                self.load(x, grid, target)
            if self.use_graph:
                (self._graphs[self._active] if self._graphs is not None else self._g_bwd).replay()
                if self._sets is not None:
                    self._set_free[self._active].record()
                if self._g_opt is not None:
                    self.grads.all_reduce_mean(self.group)
                    self._g_opt.replay()
            else:
                self._forward_backward()
                if not (self._overlap or self._ingraph or self._fused_dp):
                    self.grads.all_reduce_mean(self.group)
                self.opt.step()
        if self._loss_slots is not None:
            buf, ev = self._loss_slots[self._nsteps % 2]
            buf.copy_(self.loss, non_blocking=True)
            ev.record()
        self._nsteps += 1
        return self.loss
