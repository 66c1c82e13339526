#!/usr/bin/env python3
"""Golden vectors of the adaptive Fourier layers from the REAL reference
(/root/reference/Models/fno/spectral_layers.py:341-573), build container only:  python tests/golden/make_golden_afno.py

Imports the reference classes (nothing of the reference is copied here) and writes tests/golden/afno*.npz: input, state
dict, output and every gradient, for 1-D (relu default, odd length), 2-D (gelu, odd rows - the reference's own smoke
shape 55 x 64 scaled down - and a truncating ``hard_thresholding_fraction``) and 3-D."""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from golden_util import reference_modules  # noqa: E402


def _np(t):
    return t.detach().cpu().numpy()


def dump(name, layer, x, meta):
    y = layer(x)
    cot = torch.randn_like(y)
    (y * cot).sum().backward()
    blob = {"meta": np.array(json.dumps(meta)), "in.x": _np(x), "grad_in.x": _np(x.grad), "out": _np(y), "cot": _np(cot)}
    for k, v in layer.state_dict().items():
        blob[f"param.{k}"] = _np(v)
    for k, v in layer.named_parameters():
        blob[f"grad.{k}"] = _np(v.grad)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **blob)
    print("wrote", os.path.relpath(path), tuple(x.shape), "|y - x| / |x| =",
          float((y - x).norm() / x.norm()))


def spread(layer, gain):
    """The reference init (0.02 * U[0,1)) leaves the whole MLP output under the soft-shrink threshold; scale the
    weights and centre them so that both branches of the shrink and of the activation are exercised."""
    with torch.no_grad():
        for p in layer.parameters():
            p.mul_(gain).sub_(p.mean())


def main():
    with reference_modules() as (sl, _):
        cases = [
            ("afno1d", sl.AdaptiveFourier1d, dict(hidden_size=8, num_blocks=2), (3, 8, 15), 30.0),
            ("afno1d_frac", sl.AdaptiveFourier1d, dict(hidden_size=6, num_blocks=3, hard_thresholding_fraction=0.5,
                                                       activation="gelu", sparsity_threshold=0.05), (2, 6, 16), 40.0),
            ("afno2d", sl.AdaptiveFourier2d, dict(hidden_size=8, num_blocks=4), (2, 8, 7, 10), 30.0),
            ("afno2d_frac", sl.AdaptiveFourier2d, dict(hidden_size=4, num_blocks=1, hard_thresholding_fraction=0.1,
                                                       activation="silu"), (2, 4, 6, 9), 20.0),
            ("afno3d", sl.AdaptiveFourier3d, dict(hidden_size=4, num_blocks=2, sparsity_threshold=0.02), (2, 4, 5, 4, 6), 30.0),
        ]
        for i, (name, cls, kw, shape, gain) in enumerate(cases):
            torch.manual_seed(300 + i)
            layer = cls(**kw)
            spread(layer, gain)
            x = torch.randn(*shape, requires_grad=True)
            meta = dict(kind="afno", ndim=len(shape) - 2, modes=0, **{k: v for k, v in kw.items()})
            meta.setdefault("num_blocks", 8)
            meta.setdefault("sparsity_threshold", 0.01)
            meta.setdefault("hard_thresholding_fraction", 1)
            meta.setdefault("activation", "relu" if cls is sl.AdaptiveFourier1d else "gelu")
            dump(name, layer, x, meta)


if __name__ == "__main__":
    main()
