"""Builds and drives the host-emulated kernel library (TEST INFRASTRUCTURE, see tests/emu/cuda_emu.h).

The same capi.cu / spectral_simt.cuh that nvcc compiles for sm_100a are compiled here with g++
and -DDENO_EMULATE; "device pointers" are numpy buffers.  Used only by the CPU test-suite to
check index arithmetic, workspace carving and launch geometry before GPU time is spent.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from deno4pytorch_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC_DIR = os.path.join(ROOT, "deno4pytorch_b200", "csrc")
EMU_DIR = os.path.join(ROOT, "tests", "emu")
OUT = os.path.join(EMU_DIR, "_build", "libdeno_spectral_emu.so")

_lib = None


def _stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(SRC_DIR, f) for f in os.listdir(SRC_DIR) if f.endswith((".cu", ".cuh"))]
    deps += [os.path.join(EMU_DIR, "cuda_emu.h"), os.path.join(ROOT, "include", "deno_spectral.h"),
             os.path.join(ROOT, "include", "deno_fno.h"), os.path.join(ROOT, "include", "deno_afno.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def load():
    global _lib
    if _lib is None:
        if _stale():
            os.makedirs(os.path.dirname(OUT), exist_ok=True)
            cmd = ["g++", "-O2", "-g", "-std=c++17", "-DDENO_EMULATE", "-I", EMU_DIR, "-x", "c++",
                   os.path.join(SRC_DIR, "capi.cu"), "-shared", "-fPIC", "-o", OUT]
            subprocess.run(cmd, check=True)
        _lib = _capi.bind(C.CDLL(OUT))
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _as_blocks(weights):
    out = []
    for w in weights:
        w = np.asarray(w)
        if np.iscomplexobj(w):
            w = np.ascontiguousarray(w.astype(np.complex64))
        else:
            w = np.ascontiguousarray(w.astype(np.float32))
        out.append(w)
    return out


class EmuLayer:
    """One layer call through the emulated C ABI."""

    def __init__(self, x, weights, lin_w, lin_b, modes, act):
        self.lib = load()
        self.x = np.ascontiguousarray(x, dtype=np.float32)
        self.w = _as_blocks(weights)
        self.lin_w = np.ascontiguousarray(lin_w, dtype=np.float32)
        self.lin_b = None if lin_b is None else np.ascontiguousarray(lin_b, dtype=np.float32)
        B, cin = self.x.shape[:2]
        spatial = self.x.shape[2:]
        cout = self.lin_w.shape[0]
        modes = (modes,) * len(spatial) if isinstance(modes, int) else tuple(modes)
        self.four = len(spatial) == 4          # the 4-D stand-alone layer: deno_spectral4_* on a deno_spectral_desc4
        self.desc = _capi.make_desc4(B, cin, cout, spatial, modes, act) if self.four else \
            _capi.make_desc(B, cin, cout, spatial, modes, act)
        self.fn = (lambda name: getattr(self.lib, name.replace("deno_spectral_", "deno_spectral4_"))) if self.four else \
            (lambda name: getattr(self.lib, name))
        self.out_shape = (B, cout) + tuple(spatial)
        nb = self.fn("deno_spectral_twiddle_bytes")(C.byref(self.desc))
        if nb == 0:
            raise _capi.DenoError(self.lib.deno_last_error().decode())
        self.tw = np.zeros(nb // 4, dtype=np.float32)
        _capi.check(self.lib, self.fn("deno_spectral_twiddle_fill")(C.byref(self.desc), _ptr(self.tw), nb), "twiddle_fill")
        self.M = self.fn("deno_spectral_num_modes")(C.byref(self.desc))

    def forward(self, save=True):
        lib, d = self.lib, self.desc
        y = np.full(self.out_shape, np.nan, dtype=np.float32)
        self.z = np.full(self.out_shape, np.nan, dtype=np.float32) if save else None
        self.xhat = np.full((self.x.shape[0], self.x.shape[1], self.M), np.nan, dtype=np.complex64) if save else None
        nws = self.fn("deno_spectral_workspace_bytes")(C.byref(d), 0)
        ws = np.full(max(nws, 1), 0xFF, dtype=np.uint8)
        wptrs = _capi.pointer_array([w.ctypes.data for w in self.w])
        rc = self.fn("deno_spectral_fwd")(C.byref(d), _ptr(self.x), wptrs, _ptr(self.lin_w), _ptr(self.lin_b), _ptr(self.tw),
                                   _ptr(y), _ptr(self.z), _ptr(self.xhat), _ptr(ws), nws, None)
        _capi.check(lib, rc, "deno_spectral_fwd")
        self.fwd_launches = lib.deno_spectral_last_launch_count()
        return y

    def backward(self, gy, split_roles=False):
        """``split_roles``: through deno_spectral_bwd_overlap with an (empty) deno_overlap - the emulation has no streams,
        but the call then launches the input-gradient and parameter-gradient roles of the mode mix separately, exactly
        as the two-stream backward does on the GPU."""
        lib, d = self.lib, self.desc
        gy = np.ascontiguousarray(gy, dtype=np.float32)
        gx = np.full(self.x.shape, np.nan, dtype=np.float32)
        gws = [np.full(w.shape, np.nan, dtype=w.dtype) for w in self.w]
        glw = np.full(self.lin_w.shape, np.nan, dtype=np.float32)
        glb = np.full((self.lin_w.shape[0],), np.nan, dtype=np.float32)
        nws = self.fn("deno_spectral_workspace_bytes")(C.byref(d), 1)
        ws = np.full(max(nws, 1), 0xFF, dtype=np.uint8)
        wptrs = _capi.pointer_array([w.ctypes.data for w in self.w])
        gwptrs = _capi.pointer_array([g.ctypes.data for g in gws])
        if self.four:
            rc = lib.deno_spectral4_bwd(C.byref(d), _ptr(gy), _ptr(self.x), _ptr(self.z), _ptr(self.xhat), wptrs,
                                        _ptr(self.lin_w), _ptr(self.tw), _ptr(gx), gwptrs, _ptr(glw), _ptr(glb), _ptr(ws),
                                        nws, None, None)
        elif split_roles:
            ov = _capi.Overlap()
            _capi.check(lib, lib.deno_overlap_create(C.byref(ov)), "deno_overlap_create")
            rc = lib.deno_spectral_bwd_overlap(C.byref(d), _ptr(gy), _ptr(self.x), _ptr(self.z), _ptr(self.xhat), wptrs,
                                               _ptr(self.lin_w), _ptr(self.tw), _ptr(gx), gwptrs, _ptr(glw), _ptr(glb),
                                               _ptr(ws), nws, None, C.byref(ov))
            lib.deno_overlap_destroy(C.byref(ov))
        else:
            rc = lib.deno_spectral_bwd(C.byref(d), _ptr(gy), _ptr(self.x), _ptr(self.z), _ptr(self.xhat), wptrs,
                                       _ptr(self.lin_w), _ptr(self.tw), _ptr(gx), gwptrs, _ptr(glw), _ptr(glb), _ptr(ws),
                                       nws, None)
        _capi.check(lib, rc, "deno_spectral_bwd")
        self.bwd_launches = lib.deno_spectral_last_launch_count()
        return gx, gws, glw, glb


class EmuAfno:
    """One adaptive Fourier layer call (include/deno_afno.h) through the emulated C ABI; numpy in, numpy out."""

    def __init__(self, x, w1, b1, w2, b2, num_blocks, modes_last, act, sparsity_threshold):
        self.lib = load()
        self.x = np.ascontiguousarray(x, dtype=np.float32)
        self.p = [np.ascontiguousarray(np.asarray(a).astype(np.complex64)) for a in (w1, b1, w2, b2)]
        B, Cn = self.x.shape[:2]
        self.desc = _capi.make_afno_desc(B, Cn, num_blocks, self.x.shape[2:], modes_last, act, sparsity_threshold)
        nb = self.lib.deno_afno_twiddle_bytes(C.byref(self.desc))
        if nb == 0:
            raise _capi.DenoError(self.lib.deno_last_error().decode())
        self.tw = np.zeros(nb // 4, dtype=np.float32)
        _capi.check(self.lib, self.lib.deno_afno_twiddle_fill(C.byref(self.desc), _ptr(self.tw), nb), "deno_afno_twiddle_fill")
        assert np.array_equal(self.tw[-Cn * Cn:].reshape(Cn, Cn), np.eye(Cn, dtype=np.float32))   # the residual's bypass
        self.M = self.lib.deno_afno_num_modes(C.byref(self.desc))

    def forward(self, save=True):
        lib, d = self.lib, self.desc
        y = np.full(self.x.shape, np.nan, dtype=np.float32)
        self.xhat = np.full(self.x.shape[:2] + (self.M,), np.nan, dtype=np.complex64) if save else None
        nws = lib.deno_afno_workspace_bytes(C.byref(d), 0)
        ws = np.full(max(nws, 1), 0xFF, dtype=np.uint8)
        rc = lib.deno_afno_fwd(C.byref(d), _ptr(self.x), *[_ptr(a) for a in self.p], _ptr(self.tw), _ptr(y),
                               _ptr(self.xhat), _ptr(ws), nws, None)
        _capi.check(lib, rc, "deno_afno_fwd")
        self.fwd_launches = lib.deno_spectral_last_launch_count()
        return y

    def backward(self, gy, want_gx=True, want_params=True):
        lib, d = self.lib, self.desc
        gy = np.ascontiguousarray(gy, dtype=np.float32)
        gx = np.full(self.x.shape, np.nan, dtype=np.float32) if want_gx else None
        gp = [np.full(a.shape, np.nan, dtype=np.complex64) if want_params else None for a in self.p]
        nws = lib.deno_afno_workspace_bytes(C.byref(d), 1)
        ws = np.full(max(nws, 1), 0xFF, dtype=np.uint8)
        rc = lib.deno_afno_bwd(C.byref(d), _ptr(gy), _ptr(self.xhat), *[_ptr(a) for a in self.p], _ptr(self.tw), _ptr(gx),
                               *[_ptr(g) for g in gp], _ptr(ws), nws, None)
        _capi.check(lib, rc, "deno_afno_bwd")
        self.bwd_launches = lib.deno_spectral_last_launch_count()
        return gx, gp


class EmuEnds:
    """Lift / projection calls (include/deno_fno.h) through the emulated C ABI; numpy in, numpy out."""

    def __init__(self, batch, width, hidden, in_feats, grid_feats, out_dim, spatial, padding):
        self.lib = load()
        self.desc = _capi.make_fno_desc(batch, width, hidden, in_feats, grid_feats, out_dim, spatial, padding)
        self.spatial = tuple(spatial)
        self.padded = tuple(s + padding for s in spatial)
        self.B, self.C, self.Hd, self.O = batch, width, hidden, out_dim

    def supported(self):
        return bool(self.lib.deno_fno_supported(C.byref(self.desc)))

    def _ws(self, which):
        n = self.lib.deno_fno_workspace_bytes(C.byref(self.desc), which)
        return np.full(max(n, 1), 0xFF, dtype=np.uint8)   # NaN poison

    def lift(self, x, grid, w0, b0):
        x, grid, w0, b0 = (np.ascontiguousarray(a, dtype=np.float32) for a in (x, grid, w0, b0))
        h = np.full((self.B, self.C) + self.padded, np.nan, dtype=np.float32)
        _capi.check(self.lib, self.lib.deno_fno_lift_fwd(C.byref(self.desc), _ptr(x), _ptr(grid), _ptr(w0), _ptr(b0),
                                                         _ptr(h), None), "lift_fwd")
        return h

    def lift_bwd(self, gh, x, grid, w0, want_gx=True):
        gh, x, grid, w0 = (np.ascontiguousarray(a, dtype=np.float32) for a in (gh, x, grid, w0))
        gx = np.full(x.shape, np.nan, dtype=np.float32) if want_gx else None
        gw0 = np.full(w0.shape, np.nan, dtype=np.float32)
        gb0 = np.full(w0.shape[0], np.nan, dtype=np.float32)
        ws = self._ws(_capi.FNO_WS_LIFT_BWD)
        _capi.check(self.lib, self.lib.deno_fno_lift_bwd(C.byref(self.desc), _ptr(gh), _ptr(x), _ptr(grid), _ptr(w0),
                                                         _ptr(gx), _ptr(gw0), _ptr(gb0), _ptr(ws), ws.size, None),
                    "lift_bwd")
        return gx, gw0, gb0

    def proj(self, h, w1, b1, w2, b2):
        h, w1, b1, w2, b2 = (np.ascontiguousarray(a, dtype=np.float32) for a in (h, w1, b1, w2, b2))
        y = np.full((self.B,) + self.spatial + (self.O,), np.nan, dtype=np.float32)
        _capi.check(self.lib, self.lib.deno_fno_proj_fwd(C.byref(self.desc), _ptr(h), _ptr(w1), _ptr(b1), _ptr(w2),
                                                         _ptr(b2), _ptr(y), None), "proj_fwd")
        return y

    def proj_bwd(self, gy, h, w1, b1, w2):
        gy, h, w1, b1, w2 = (np.ascontiguousarray(a, dtype=np.float32) for a in (gy, h, w1, b1, w2))
        gh = np.full(h.shape, np.nan, dtype=np.float32)
        gw1 = np.full(w1.shape, np.nan, dtype=np.float32)
        gb1 = np.full(b1.shape, np.nan, dtype=np.float32)
        gw2 = np.full(w2.shape, np.nan, dtype=np.float32)
        gb2 = np.full(w2.shape[0], np.nan, dtype=np.float32)
        ws = self._ws(_capi.FNO_WS_PROJ_BWD)
        _capi.check(self.lib, self.lib.deno_fno_proj_bwd(C.byref(self.desc), _ptr(gy), _ptr(h), _ptr(w1), _ptr(b1),
                                                         _ptr(w2), _ptr(gh), _ptr(gw1), _ptr(gb1), _ptr(gw2), _ptr(gb2),
                                                         _ptr(ws), ws.size, None), "proj_bwd")
        return gh, gw1, gb1, gw2, gb2
