"""Loader for the committed reference fixtures (tests/golden/*.npz)."""
import glob
import json
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names(kind=None):
    names = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    if kind is None:
        return names
    return [n for n in names if load_golden(n)["meta"]["kind"] == kind]


def load_golden(name):
    d = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    out = {"meta": json.loads(str(d["meta"])), "inputs": {}, "grad_inputs": {}, "params": {}, "grads": {}}
    for k in d.files:
        if k == "meta":
            continue
        t = torch.from_numpy(d[k])
        if k.startswith("in."):
            out["inputs"][k[3:]] = t
        elif k.startswith("grad_in."):
            out["grad_inputs"][k[8:]] = t
        elif k.startswith("param."):
            out["params"][k[6:]] = t
        elif k.startswith("grad."):
            out["grads"][k[5:]] = t
        else:
            out[k] = t
    m = out["meta"]["modes"]
    out["meta"]["modes"] = tuple(m) if isinstance(m, list) else m
    return out


def rel_l2(a, b):
    a = torch.as_tensor(a)
    b = torch.as_tensor(b)
    if a.is_complex() != b.is_complex():
        a = torch.view_as_real(a) if a.is_complex() else a
        b = torch.view_as_real(b) if b.is_complex() else b
    a = a.to(torch.complex128 if a.is_complex() else torch.float64)
    b = b.to(a.dtype)
    denom = torch.linalg.vector_norm(b).item()
    num = torch.linalg.vector_norm(a - b).item()
    return num / max(denom, 1e-300)


import contextlib
import sys

REFERENCE_ROOT = os.environ.get("DENO_REFERENCE", "/root/reference")
_REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@contextlib.contextmanager
def reference_modules():
    """Imports the REAL reference ``Models.fno`` (build container only).  This repo ships regular
    packages called ``Models`` / ``fno`` (the drop-in), which would shadow the reference's
    namespace packages, so they are hidden from sys.path / sys.modules for the duration."""
    def ours(name):
        return name in ("Models", "fno") or name.startswith(("Models.", "fno."))

    saved_modules = {k: sys.modules.pop(k) for k in list(sys.modules) if ours(k)}
    saved_path = list(sys.path)
    saved_cwd = os.getcwd()
    hidden = {_REPO_ROOT, os.path.join(_REPO_ROOT, "Models"), ""}
    sys.path[:] = [REFERENCE_ROOT] + [p for p in saved_path if os.path.abspath(p or saved_cwd) not in hidden and p != ""]
    os.chdir("/tmp")
    try:
        from Models.fno import FNOs, spectral_layers
        assert spectral_layers.__file__.startswith(REFERENCE_ROOT), spectral_layers.__file__
        yield spectral_layers, FNOs
    finally:
        os.chdir(saved_cwd)
        sys.path[:] = saved_path
        for k in [k for k in sys.modules if ours(k)]:
            del sys.modules[k]
        sys.modules.update(saved_modules)
