"""Synthetic inputs of the demo shapes (no dataset ships with the reference):
``x ~ N(0,1)``, ``target ~ N(0,1)``, ``grid`` = the ``linspace(0,1)`` coordinate mesh every demo's
``feature_transform`` appends (/root/reference/Demo/Darcy_2d/run_FNO&UNet.py:27-40)."""
from __future__ import annotations

from typing import Sequence

import torch


def make_grid(batch: int, spatial: Sequence[int]) -> torch.Tensor:
    axes = [torch.linspace(0, 1, n, dtype=torch.float32) for n in spatial]
    mesh = torch.stack(torch.meshgrid(*axes, indexing="ij"), dim=-1)
    return mesh.unsqueeze(0).expand(batch, *mesh.shape).contiguous()


def synthetic_batch(ndim: int, model_kwargs: dict, spatial: Sequence[int], batch: int, seed: int = 0):
    """Returns CPU tensors ``(x, grid, target)`` shaped for ``FNO{ndim}d(**model_kwargs)``."""
    assert len(spatial) == ndim
    gen = torch.Generator().manual_seed(seed)
    cin = model_kwargs["in_dim"] * model_kwargs.get("steps", 1)
    x = torch.randn((batch,) + tuple(spatial) + (cin,), generator=gen)
    target = torch.randn((batch,) + tuple(spatial) + (model_kwargs["out_dim"],), generator=gen)
    return x, make_grid(batch, spatial), target
