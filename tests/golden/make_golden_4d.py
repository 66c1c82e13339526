#!/usr/bin/env python3
"""Golden vectors of the 4-D stand-alone layer / model from the REAL reference
(/root/reference/Demo/Turbulence_3d+t/FNO_HIT.py:31-171), build container only:  python tests/golden/make_golden_4d.py

FNO_HIT.py is a script (it imports torchvision / matplotlib / a local ``utilities3`` and then trains), so it cannot be
imported.  The two class definitions are taken from its syntax tree AT RUN TIME and executed unmodified in a namespace
that holds what they use (torch, nn, F, np); nothing of the reference is copied into this repository.
Writes tests/golden/layer4d.npz and tests/golden/fno4d.npz (inputs, state dict, output, every gradient)."""
import ast
import json
import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

REF = os.environ.get("DENO_REFERENCE", "/root/reference")
SRC = os.path.join(REF, "Demo", "Turbulence_3d+t", "FNO_HIT.py")
HERE = os.path.dirname(os.path.abspath(__file__))


def reference_classes():
    with open(SRC) as f:
        tree = ast.parse(f.read(), SRC)
    wanted = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name in ("SpectralConv4d", "FNO4d")]
    assert [n.name for n in wanted] == ["SpectralConv4d", "FNO4d"]
    ns = {"torch": torch, "nn": nn, "F": F, "np": np}
    exec(compile(ast.Module(body=wanted, type_ignores=[]), SRC, "exec"), ns)
    return ns["SpectralConv4d"], ns["FNO4d"]


def _np(t):
    return t.detach().cpu().numpy()


def dump(path, module, inputs, out, cot, meta):
    (out * cot).sum().backward()
    blob = {"meta": np.array(json.dumps(meta))}
    for k, v in inputs.items():
        blob[f"in.{k}"] = _np(v)
        blob[f"grad_in.{k}"] = _np(v.grad)
    for k, v in module.state_dict().items():
        blob[f"param.{k}"] = _np(v)
    for k, v in module.named_parameters():
        blob[f"grad.{k}"] = _np(v.grad)
    blob["out"] = _np(out)
    blob["cot"] = _np(cot)
    np.savez_compressed(path, **blob)
    print("wrote", os.path.relpath(path), {k: v.shape for k, v in blob.items() if k.startswith("in.")})


def main():
    if not os.path.exists(SRC):
        raise SystemExit(f"reference script not found at {SRC}")
    SpectralConv4d, FNO4d = reference_classes()
    # layer: odd / even sizes, the last axis has the reference's length 3 (modes4 is clamped to 2 by the constructor)
    torch.manual_seed(77)
    cin, cout, modes = 3, 2, (2, 3, 2, 2)
    layer = SpectralConv4d(cin, cout, *modes)
    with torch.no_grad():
        for p in layer.parameters():
            p.mul_(cin * cout * 2).sub_(p.mean())      # O(1) spectral branch, both signs in real and imaginary parts
    x = torch.randn(2, cin, 6, 7, 5, 3, requires_grad=True)
    y = layer(x)
    dump(os.path.join(HERE, "layer4d.npz"), layer, {"x": x}, y, torch.randn_like(y),
         dict(kind="layer4d", ndim=4, in_dim=cin, out_dim=cout, modes=[layer.modes1, layer.modes2, layer.modes3, layer.modes4]))
    # model: FNO4d(modes1..4, width); input (B, X, Y, Z, T, 5), grid appended inside
    torch.manual_seed(78)
    width, mm = 4, (2, 2, 2, 2)
    model = FNO4d(*mm, width)
    with torch.no_grad():
        for k, p in model.named_parameters():
            if ".weights" in k:
                p.mul_(width ** 2 / 2).sub_(p.mean())
    x = torch.randn(2, 5, 4, 6, 3, 5, requires_grad=True)
    y = model(x)
    dump(os.path.join(HERE, "fno4d.npz"), model, {"x": x}, y, torch.randn_like(y),
         dict(kind="model4d", ndim=4, width=width, modes=[model.conv0.modes1, model.conv0.modes2, model.conv0.modes3,
                                                          model.conv0.modes4]))


if __name__ == "__main__":
    main()
