#!/usr/bin/env python3
"""Generate golden input/output vectors from the REAL reference.

Run in the build container only (``/root/reference`` is mounted there and nowhere
else):  ``python tests/golden/make_golden.py``.  It imports the unmodified
reference classes (Models/fno/spectral_layers.py, Models/fno/FNOs.py), runs them
on seeded CPU inputs and writes ``tests/golden/*.npz`` holding inputs, the full
state dict, outputs and every gradient.  The fixtures are committed; this script
is committed beside them so they can be regenerated and audited.

Loss used for gradients: ``sum(y * r)`` with a stored random cotangent ``r``.
"""
import json
import os
import sys

import numpy as np
import torch

REF = os.environ.get("DENO_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def _import_reference():
    if not os.path.isdir(REF):
        raise SystemExit(f"reference tree not found at {REF}")
    sys.path.insert(0, os.path.dirname(HERE))
    from golden_util import reference_modules
    ctx = reference_modules()
    mods = ctx.__enter__()  # kept for the life of the script: the reference must win over the drop-in
    return mods


def _np(t):
    return t.detach().cpu().numpy()


def _dump(path, module, inputs, out, cot, meta):
    (out * cot).sum().backward()
    blob = {"meta": np.array(json.dumps(meta))}
    for k, v in inputs.items():
        blob[f"in.{k}"] = _np(v)
        if v.grad is not None:
            blob[f"grad_in.{k}"] = _np(v.grad)
    for k, v in module.state_dict().items():
        blob[f"param.{k}"] = _np(v)
    for k, v in module.named_parameters():
        blob[f"grad.{k}"] = _np(v.grad)
    blob["out"] = _np(out)
    blob["cot"] = _np(cot)
    np.savez_compressed(path, **blob)
    print("wrote", os.path.relpath(path), {k: v.shape for k, v in blob.items() if k.startswith("in.")})


LAYER_CASES = [
    # name, ndim, in, out, modes, spatial, batch, activation, norm, use_complex, weight_gain
    ("layer1d_c", 1, 3, 5, 4, (16,), 3, "gelu", None, True, 1.0),
    ("layer1d_r", 1, 4, 4, 5, (18,), 2, "silu", "ortho", False, 16.0),
    ("layer1d_nyq", 1, 2, 3, 7, (12,), 2, "tanh", None, True, 6.0),
    ("layer2d_c", 2, 3, 4, (3, 4), (10, 12), 2, "gelu", None, True, 12.0),
    ("layer2d_odd", 2, 4, 4, (2, 3), (9, 11), 2, "gelu", None, True, 16.0),
    ("layer2d_ortho", 2, 2, 3, 3, (8, 10), 2, "silu", "ortho", True, 6.0),
    ("layer2d_nyq", 2, 3, 3, (2, 5), (7, 8), 2, "relu", "forward", True, 9.0),
    ("layer3d_r", 3, 3, 2, (2, 2, 3), (8, 6, 10), 2, "gelu", None, False, 6.0),
    ("layer3d_r2", 3, 4, 4, 2, (5, 7, 6), 2, "leakyrelu", None, False, 16.0),
]

MODEL_CASES = [
    # name, ndim, ctor kwargs, spatial, batch
    ("fno1d", 1, dict(in_dim=1, out_dim=1, modes=4, width=8, depth=2, hidden=16, steps=1, padding=2), (14,), 3),
    ("fno1d_r", 1, dict(in_dim=2, out_dim=2, modes=3, width=6, depth=2, hidden=12, steps=1, padding=0,
                        use_complex=False), (12,), 2),
    ("fno2d", 2, dict(in_dim=1, out_dim=1, modes=(3, 3), width=8, depth=2, hidden=16, steps=1, padding=3), (9, 9), 2),
    ("fno2d_nopad", 2, dict(in_dim=3, out_dim=2, modes=(2, 3), width=6, depth=3, hidden=10, steps=1, padding=0),
     (8, 10), 2),
    ("fno3d", 3, dict(in_dim=2, out_dim=1, modes=(2, 2, 2), width=6, depth=2, hidden=12, steps=1, padding=2,
                      use_complex=False), (6, 5, 4), 2),
]


def main():
    spectral_layers, FNOs = _import_reference()
    layer_cls = {1: spectral_layers.SpectralConv1d, 2: spectral_layers.SpectralConv2d,
                 3: spectral_layers.SpectralConv3d}
    model_cls = {1: FNOs.FNO1d, 2: FNOs.FNO2d, 3: FNOs.FNO3d}

    for (name, nd, cin, cout, modes, spatial, batch, act, norm, use_complex, gain) in LAYER_CASES:
        torch.manual_seed(1234 + len(name))
        layer = layer_cls[nd](cin, cout, modes, dropout=0.0, norm=norm, activation=act, use_complex=use_complex)
        with torch.no_grad():
            for k, p in layer.named_parameters():
                if k.startswith("weights"):
                    # lift the spectral branch to O(1) of the bypass (SURVEY section 4 hazard) and
                    # centre it so real/imag parts carry both signs
                    p.mul_(gain * cin * cout / 4).sub_(p.mean())
        x = torch.randn(batch, cin, *spatial, requires_grad=True)
        y = layer(x)
        cot = torch.randn_like(y)
        meta = dict(kind="layer", ndim=nd, in_dim=cin, out_dim=cout, modes=modes, activation=act, norm=norm,
                    use_complex=use_complex)
        _dump(os.path.join(HERE, f"{name}.npz"), layer, {"x": x}, y, cot, meta)

    for (name, nd, kw, spatial, batch) in MODEL_CASES:
        torch.manual_seed(4321 + len(name))
        model = model_cls[nd](**kw)
        with torch.no_grad():
            for k, p in model.named_parameters():
                if ".weights" in k:
                    p.mul_(kw["width"] ** 2 / 2).sub_(p.mean())
        x = torch.randn(batch, *spatial, kw["in_dim"] * kw["steps"], requires_grad=True)
        grid = torch.rand(batch, *spatial, nd)
        y = model(x, grid)
        cot = torch.randn_like(y)
        meta = dict(kind="model", ndim=nd, **kw)
        _dump(os.path.join(HERE, f"{name}.npz"), model, {"x": x, "grid": grid}, y, cot, meta)


if __name__ == "__main__":
    main()
