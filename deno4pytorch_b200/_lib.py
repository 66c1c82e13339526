"""Loader of the CUDA library.  There is NO CPU fallback: if ``libdeno_spectral.so`` has not been
built (``python -c "import __graft_entry__ as g; g.build()"``) every entry point raises."""
from __future__ import annotations

import ctypes
import os

from . import _capi

_HERE = os.path.dirname(os.path.abspath(__file__))
# DENO_LIB_PATH: A/B timing of an older build of the same ABI (tools/ab_layer.py); never a different implementation
LIB_PATH = os.environ.get("DENO_LIB_PATH") or os.path.join(_HERE, "csrc", "libdeno_spectral.so")
_lib = None


class LibraryMissing(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LibraryMissing(
                f"{LIB_PATH} not found - the sm_100a CUDA extension has not been built "
                "(run `python -c \"import __graft_entry__ as g; g.build()\"`). "
                "deno4pytorch_b200 has no CPU or PyTorch fallback.")
        _lib = _capi.bind(ctypes.CDLL(LIB_PATH))
    return _lib
