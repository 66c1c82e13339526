"""ctypes view of include/deno_spectral.h (no torch imports here).

``bind(cdll)`` attaches argument / return types for every exported symbol and returns the
library.  ``EXPORTED_SYMBOLS`` is the list the CPU test-suite checks the built ``.so`` against.
"""
from __future__ import annotations

import ctypes as C

ABI_VERSION = 1

ACTIVATION_IDS = {"identity": 0, "gelu": 1, "silu": 2, "relu": 3, "tanh": 4, "leakyrelu": 5}

STATUS_NAMES = {0: "DENO_OK", 1: "DENO_E_INVALID", 2: "DENO_E_UNSUPPORTED", 3: "DENO_E_WORKSPACE", 4: "DENO_E_CUDA"}


class SpectralDesc(C.Structure):
    """struct deno_spectral_desc"""
    _fields_ = [
        ("ndim", C.c_int32),
        ("batch", C.c_int32),
        ("cin", C.c_int32),
        ("cout", C.c_int32),
        ("spatial", C.c_int32 * 3),
        ("modes", C.c_int32 * 3),
        ("act", C.c_int32),
        ("flags", C.c_int32),
    ]


def make_desc(batch, cin, cout, spatial, modes, act) -> SpectralDesc:
    nd = len(spatial)
    assert len(modes) == nd and 1 <= nd <= 3
    d = SpectralDesc()
    d.ndim, d.batch, d.cin, d.cout = nd, int(batch), int(cin), int(cout)
    for a in range(3):
        d.spatial[a] = int(spatial[a]) if a < nd else 0
        d.modes[a] = int(modes[a]) if a < nd else 0
    d.act = ACTIVATION_IDS[act] if isinstance(act, str) else int(act)
    d.flags = 0
    return d


class SpectralDesc4(C.Structure):
    """struct deno_spectral_desc4 (the 4-D stand-alone layer of Demo/Turbulence_3d+t/FNO_HIT.py)"""
    _fields_ = [
        ("batch", C.c_int32),
        ("cin", C.c_int32),
        ("cout", C.c_int32),
        ("spatial", C.c_int32 * 4),
        ("modes", C.c_int32 * 4),
        ("act", C.c_int32),
        ("flags", C.c_int32),
    ]
    ndim = 4


def make_desc4(batch, cin, cout, spatial, modes, act) -> SpectralDesc4:
    assert len(spatial) == 4 and len(modes) == 4
    d = SpectralDesc4()
    d.batch, d.cin, d.cout = int(batch), int(cin), int(cout)
    for a in range(4):
        d.spatial[a] = int(spatial[a])
        d.modes[a] = int(modes[a])
    d.act = ACTIVATION_IDS[act] if isinstance(act, str) else int(act)
    d.flags = 0
    return d


class AfnoDesc(C.Structure):
    """struct deno_afno_desc (include/deno_afno.h)"""
    _fields_ = [
        ("ndim", C.c_int32),
        ("batch", C.c_int32),
        ("channels", C.c_int32),
        ("num_blocks", C.c_int32),
        ("spatial", C.c_int32 * 3),
        ("modes_last", C.c_int32),
        ("act", C.c_int32),
        ("sparsity_threshold", C.c_float),
        ("flags", C.c_int32),
    ]


def make_afno_desc(batch, channels, num_blocks, spatial, modes_last, act, sparsity_threshold) -> AfnoDesc:
    nd = len(spatial)
    assert 1 <= nd <= 3
    d = AfnoDesc()
    d.ndim, d.batch, d.channels, d.num_blocks = nd, int(batch), int(channels), int(num_blocks)
    for a in range(3):
        d.spatial[a] = int(spatial[a]) if a < nd else 0
    d.modes_last = int(modes_last)
    d.act = ACTIVATION_IDS[act] if isinstance(act, str) else int(act)
    d.sparsity_threshold = float(sparsity_threshold)
    d.flags = 0
    return d


class FnoDesc(C.Structure):
    """struct deno_fno_desc (include/deno_fno.h)"""
    _fields_ = [
        ("ndim", C.c_int32),
        ("batch", C.c_int32),
        ("width", C.c_int32),
        ("hidden", C.c_int32),
        ("in_feats", C.c_int32),
        ("grid_feats", C.c_int32),
        ("out_dim", C.c_int32),
        ("spatial", C.c_int32 * 3),
        ("padding", C.c_int32),
    ]


FNO_WS_LIFT_BWD, FNO_WS_PROJ_BWD = 0, 1


def make_fno_desc(batch, width, hidden, in_feats, grid_feats, out_dim, spatial, padding) -> FnoDesc:
    nd = len(spatial)
    assert 1 <= nd <= 3
    d = FnoDesc()
    d.ndim, d.batch, d.width, d.hidden = nd, int(batch), int(width), int(hidden)
    d.in_feats, d.grid_feats, d.out_dim, d.padding = int(in_feats), int(grid_feats), int(out_dim), int(padding)
    for a in range(3):
        d.spatial[a] = int(spatial[a]) if a < nd else 0
    return d


class Overlap(C.Structure):
    """struct deno_overlap: side stream + fork / join events of deno_spectral_bwd_overlap"""
    _fields_ = [("side_stream", C.c_void_p), ("ev_fork", C.c_void_p), ("ev_join", C.c_void_p)]


class ProfileRecord(C.Structure):
    """struct deno_profile_record"""
    _fields_ = [("name", C.c_char * 32), ("ms", C.c_float)]


_DESC_P = C.POINTER(SpectralDesc)
_DESC4_P = C.POINTER(SpectralDesc4)
_FNO_P = C.POINTER(FnoDesc)
_AFNO_P = C.POINTER(AfnoDesc)
_V = C.c_void_p
_PP = C.POINTER(C.c_void_p)

_PROTOTYPES = {
    "deno_spectral_abi_version": (C.c_int, []),
    "deno_last_error": (C.c_char_p, []),
    "deno_reload_env": (C.c_int, []),
    "deno_spectral_last_launch_count": (C.c_int, []),
    "deno_profile_begin": (C.c_int, []),
    "deno_profile_end": (C.c_int, [C.POINTER(ProfileRecord), C.c_int]),
    "deno_debug_read": (C.c_int, [C.POINTER(C.c_longlong), C.c_int, C.c_int]),
    "deno_adam_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_float, C.c_float,
                                 C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p]),
    "deno_dp_fused_adam": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int,
                                     C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                     C.c_void_p, C.c_void_p]),
    "deno_spectral_num_modes": (C.c_size_t, [_DESC_P]),
    "deno_spectral_twiddle_bytes": (C.c_size_t, [_DESC_P]),
    "deno_spectral_twiddle_fill": (C.c_int, [_DESC_P, C.c_void_p, C.c_size_t]),
    "deno_spectral_workspace_bytes": (C.c_size_t, [_DESC_P, C.c_int]),
    "deno_spectral_xhat_bytes": (C.c_size_t, [_DESC_P]),
    "deno_spectral_fwd": (C.c_int, [_DESC_P, C.c_void_p, _PP, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "deno_spectral_bwd": (C.c_int, [_DESC_P, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, _PP, C.c_void_p,
                                    C.c_void_p, C.c_void_p, _PP, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                    C.c_void_p]),
    "deno_overlap_create": (C.c_int, [C.POINTER(Overlap)]),
    "deno_overlap_destroy": (C.c_int, [C.POINTER(Overlap)]),
    "deno_spectral_bwd_overlap": (C.c_int, [_DESC_P, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, _PP, C.c_void_p,
                                            C.c_void_p, C.c_void_p, _PP, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                            C.c_void_p, C.POINTER(Overlap)]),
    "deno_spectral4_num_modes": (C.c_size_t, [_DESC4_P]),
    "deno_spectral4_twiddle_bytes": (C.c_size_t, [_DESC4_P]),
    "deno_spectral4_twiddle_fill": (C.c_int, [_DESC4_P, C.c_void_p, C.c_size_t]),
    "deno_spectral4_workspace_bytes": (C.c_size_t, [_DESC4_P, C.c_int]),
    "deno_spectral4_fwd": (C.c_int, [_DESC4_P, C.c_void_p, _PP, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "deno_spectral4_bwd": (C.c_int, [_DESC4_P, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, _PP, C.c_void_p,
                                     C.c_void_p, C.c_void_p, _PP, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                     C.c_void_p, C.POINTER(Overlap)]),
    "deno_afno_num_modes": (C.c_size_t, [_AFNO_P]),
    "deno_afno_twiddle_bytes": (C.c_size_t, [_AFNO_P]),
    "deno_afno_twiddle_fill": (C.c_int, [_AFNO_P, C.c_void_p, C.c_size_t]),
    "deno_afno_workspace_bytes": (C.c_size_t, [_AFNO_P, C.c_int]),
    "deno_afno_fwd": (C.c_int, [_AFNO_P, _V, _V, _V, _V, _V, _V, _V, _V, _V, C.c_size_t, _V]),
    "deno_afno_bwd": (C.c_int, [_AFNO_P, _V, _V, _V, _V, _V, _V, _V, _V, _V, _V, _V, _V, _V, C.c_size_t, _V]),
    "deno_fno_supported": (C.c_int, [_FNO_P]),
    "deno_fno_workspace_bytes": (C.c_size_t, [_FNO_P, C.c_int]),
    "deno_fno_lift_fwd": (C.c_int, [_FNO_P, _V, _V, _V, _V, _V, _V]),
    "deno_fno_lift_bwd": (C.c_int, [_FNO_P, _V, _V, _V, _V, _V, _V, _V, _V, C.c_size_t, _V]),
    "deno_fno_proj_fwd": (C.c_int, [_FNO_P, _V, _V, _V, _V, _V, _V, _V]),
    "deno_fno_proj_bwd": (C.c_int, [_FNO_P, _V, _V, _V, _V, _V, _V, _V, _V, _V, _V, _V, C.c_size_t, _V]),
}

EXPORTED_SYMBOLS = tuple(_PROTOTYPES)


def bind(lib: C.CDLL) -> C.CDLL:
    for name, (res, args) in _PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    got = lib.deno_spectral_abi_version()
    if got != ABI_VERSION:
        raise RuntimeError(f"libdeno_spectral ABI version {got}, expected {ABI_VERSION}")
    return lib


def pointer_array(ptrs):
    """HOST array of device pointers (w_blocks / gw_blocks)."""
    arr = (C.c_void_p * 4)()
    for j in range(4):
        arr[j] = ptrs[j] if j < len(ptrs) else None
    return arr


class DenoError(RuntimeError):
    pass


def check(lib: C.CDLL, rc: int, what: str):
    if rc != 0:
        msg = lib.deno_last_error().decode("utf-8", "replace")
        raise DenoError(f"{what} failed with {STATUS_NAMES.get(rc, rc)}: {msg}")
