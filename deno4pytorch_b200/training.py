"""``TrainStep`` - one optimiser step of an FNO stack, the way every reference demo trains
(forward, MSE, backward, Adam: /root/reference/Demo/Darcy_2d/run_FNO&UNet.py:43-69,206-209),
captured once into a CUDA graph and replayed.

The reference loop is launch-bound on small grids (each step issues a few hundred tiny library
kernels from Python); the captured graph removes the per-launch host cost.  Under data
parallelism the gradient all-reduce (deno4pytorch_b200.parallel.FlatGradients) runs between the
backward graph and the optimiser graph on the same stream.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.distributed as dist
import torch.nn.functional as F

from .parallel import FlatAdam, FlatGradients, FlatParameters


class TrainStep:
    def __init__(self, model: torch.nn.Module, x: torch.Tensor, grid: torch.Tensor, target: torch.Tensor, *,
                 lr: float = 1e-3, betas=(0.7, 0.9), weight_decay: float = 1e-4,
                 loss_fn: Callable = F.mse_loss, forward_fn: Optional[Callable] = None,
                 use_graph: bool = True, warmup: int = 3, group=None, fused_adam: bool = True):
        """``x``/``grid``/``target`` are example device tensors fixing the shapes; later calls copy new
        data into static buffers.  ``forward_fn(model, x, grid, target) -> loss`` overrides the default
        single model call (e.g. the 10-step autoregressive rollout of Demo/Turbulence_2d+t)."""
        self.model = model
        self.loss_fn = loss_fn
        self.forward_fn = forward_fn
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.sx, self.sg, self.st = x.clone(), grid.clone(), target.clone()
        if fused_adam and x.is_cuda:
            self.flat_params = FlatParameters(model.parameters())
            self.grads = self.flat_params.grads
            self.opt = FlatAdam(self.flat_params, lr=lr, betas=betas, weight_decay=weight_decay)
        else:
            self.flat_params = None
            self.grads = FlatGradients(model.parameters())
            self.opt = torch.optim.Adam(model.parameters(), lr=lr, betas=betas, weight_decay=weight_decay,
                                        capturable=use_graph, foreach=True)
        # one model call per step: parameter gradients go straight into the flat buffer (no per-parameter add kernels)
        self.grads.direct_writes(forward_fn is None and x.is_cuda)
        self.loss = torch.zeros((), device=x.device)
        self.use_graph = use_graph
        self._g_bwd = self._g_opt = None
        if use_graph:
            self._capture(warmup)

    # ------------------------------------------------------------------------------------
    def _forward_backward(self):
        self.grads.zero()
        if self.forward_fn is not None:
            loss = self.forward_fn(self.model, self.sx, self.sg, self.st)
        else:
            loss = self.loss_fn(self.model(self.sx, self.sg), self.st)
        loss.backward()
        self.loss.copy_(loss.detach())

    def _capture(self, warmup: int):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):   # also fills the twiddle cache and cuBLAS workspaces
                self._forward_backward()
                self.grads.all_reduce_mean(self.group)
                self.opt.step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._g_bwd = torch.cuda.CUDAGraph()
        if self.world == 1:
            # no collective between backward and optimiser: the whole step is one graph launch
            with torch.cuda.graph(self._g_bwd):
                self._forward_backward()
                self.opt.step()
            return
        with torch.cuda.graph(self._g_bwd):
            self._forward_backward()
        self._g_opt = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._g_opt):
            self.opt.step()

    # ------------------------------------------------------------------------------------
    def load(self, x=None, grid=None, target=None, non_blocking: bool = True):
        """Copy a new batch (host or device tensors) into the static buffers."""
        if x is not None:
            self.sx.copy_(x, non_blocking=non_blocking)
        if grid is not None:
            self.sg.copy_(grid, non_blocking=non_blocking)
        if target is not None:
            self.st.copy_(target, non_blocking=non_blocking)

    def __call__(self, x=None, grid=None, target=None) -> torch.Tensor:
        """Runs one optimiser step; returns the (device) loss of this rank's shard."""
        self.load(x, grid, target)
        if self.use_graph:
            self._g_bwd.replay()
            if self._g_opt is not None:
                self.grads.all_reduce_mean(self.group)
                self._g_opt.replay()
        else:
            self._forward_backward()
            self.grads.all_reduce_mean(self.group)
            self.opt.step()
        return self.loss
