"""Drop-in for the 4-D stand-alone model of the reference, ``SpectralConv4d`` / ``FNO4d``
(/root/reference/Demo/Turbulence_3d+t/FNO_HIT.py:31-171), on the CUDA kernels of this package.

Same constructor signatures, attribute names and state-dict keys (``conv{i}.weights{1..4}``, ``w{i}.weight/bias``,
``fc0/fc1/fc2``) as the reference script's classes, so its checkpoints load unchanged.  One model step
``conv_i(x) + w_i(x)`` (+ GELU) of ``FNO4d.forward`` (:131-150) is ONE fused layer call (``deno_spectral4_fwd``: truncated
4-D transform, four corner blocks, zero-padded inverse, 1x1 bypass, bias, activation); the bare ``SpectralConv4d.forward``
is the same call with a zero bypass.  The point-wise ends (``fc0``; ``fc1 -> GELU -> fc2`` with 512 hidden units) stay on
torch CUDA ops.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .functional import spectral_conv


class SpectralConv4d(nn.Module):
    def __init__(self, in_channels, out_channels, modes1, modes2, modes3, modes4):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.modes1 = modes1
        self.modes2 = modes2
        self.modes3 = modes3
        self.modes4 = min(modes4, 3 // 2 + 1)      # the reference clamps to the rfft length of its 3-step time axis (:44)
        self.scale = 1 / (in_channels * out_channels)
        shape = (in_channels, out_channels, self.modes1, self.modes2, self.modes3, self.modes4)
        for j in (1, 2, 3, 4):
            setattr(self, f"weights{j}", nn.Parameter(self.scale * torch.rand(*shape, dtype=torch.cfloat)))
        self.register_buffer("_zero_bypass", torch.zeros(out_channels, in_channels), persistent=False)

    @property
    def modes(self):
        return (self.modes1, self.modes2, self.modes3, self.modes4)

    def blocks(self):
        return [self.weights1, self.weights2, self.weights3, self.weights4]

    def forward(self, x):
        return spectral_conv(x, self.blocks(), self._zero_bypass, None, self.modes, "identity")


class FNO4d(nn.Module):
    def __init__(self, modes1, modes2, modes3, modes4, width):
        super().__init__()
        self.modes1, self.modes2, self.modes3, self.modes4 = modes1, modes2, modes3, modes4
        self.width = width
        self.fc0 = nn.Linear(9, self.width)        # 5 time steps + 4 coordinates
        for i in range(4):
            setattr(self, f"conv{i}", SpectralConv4d(self.width, self.width, modes1, modes2, modes3, modes4))
        for i in range(4):
            setattr(self, f"w{i}", nn.Conv1d(self.width, self.width, 1))
        self.fc1 = nn.Linear(self.width, 512)
        self.fc2 = nn.Linear(512, 1)

    def forward(self, x):
        batchsize = x.shape[0]
        sizes = tuple(x.shape[1:5])
        grid = self.get_grid(batchsize, *sizes, x.device)
        x = torch.cat((x, grid), dim=-1)
        x = self.fc0(x)
        x = x.permute(0, 5, 1, 2, 3, 4).contiguous()
        for i in range(4):
            conv, w = getattr(self, f"conv{i}"), getattr(self, f"w{i}")
            # x1 = conv(x); x2 = w(x); x = gelu(x1 + x2) (no activation after the last one): one fused call
            x = spectral_conv(x, conv.blocks(), w.weight, w.bias, conv.modes, "gelu" if i < 3 else "identity")
        x = x.permute(0, 2, 3, 4, 5, 1)
        x = F.gelu(self.fc1(x))
        return self.fc2(x)

    def get_grid(self, batchsize, size_x, size_y, size_z, size_w, device):
        axes = [torch.linspace(0, 1, n, dtype=torch.float32, device=device) for n in (size_x, size_y, size_z, size_w)]
        mesh = torch.stack(torch.meshgrid(*axes, indexing="ij"), dim=-1)
        return mesh.unsqueeze(0).expand(batchsize, *mesh.shape)
