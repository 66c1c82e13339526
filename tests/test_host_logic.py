"""CPU-only tests of the host side: C-ABI exports, drop-in module structure, import aliases,
data-parallel gradient plumbing (gloo, world size 2).  No kernel is executed here."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

from golden_util import golden_names, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    from deno4pytorch_b200 import build
    return build.build()


def _header_functions():
    text = "".join(open(os.path.join(ROOT, "include", h)).read() for h in ("deno_spectral.h", "deno_fno.h", "deno_dp.h", "deno_afno.h"))
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(deno_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built_lib):
    from deno4pytorch_b200 import _capi
    declared = _header_functions()
    assert declared == sorted(_capi.EXPORTED_SYMBOLS)
    lib = ctypes.CDLL(built_lib)
    for name in declared:
        assert hasattr(lib, name), name
    _capi.bind(lib)  # also checks the ABI version
    nm = subprocess.run(["nm", "-D", "--defined-only", built_lib], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (deno_\w+)", nm))
    assert exported == set(declared)


def test_library_is_sm100a_sass(built_lib):
    out = subprocess.run(["cuobjdump", "-lelf", built_lib], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_host_only_queries_work_without_gpu(built_lib):
    """Shape bookkeeping entry points are pure host code (no compute call is made here)."""
    from deno4pytorch_b200 import _capi
    lib = _capi.bind(ctypes.CDLL(built_lib))
    d = _capi.make_desc(20, 32, 32, (94, 94), (12, 12), "gelu")
    assert lib.deno_spectral_num_modes(ctypes.byref(d)) == 288
    assert lib.deno_spectral_xhat_bytes(ctypes.byref(d)) == 20 * 32 * 288 * 8
    assert lib.deno_spectral_workspace_bytes(ctypes.byref(d), 0) >= 2 * 20 * 32 * 288 * 8
    assert lib.deno_spectral_workspace_bytes(ctypes.byref(d), 1) >= 20 * 32 * 94 * 94 * 4
    d3 = _capi.make_desc(10, 32, 32, (70, 70, 46), (8, 8, 8), "gelu")
    assert lib.deno_spectral_num_modes(ctypes.byref(d3)) == 2048
    bad = _capi.make_desc(1, 2, 2, (15,), (3,), "gelu")
    assert lib.deno_spectral_twiddle_bytes(ctypes.byref(bad)) == 0
    assert b"even length" in lib.deno_last_error()


def test_twiddle_tables_are_fp64_accurate(built_lib):
    import numpy as np
    from deno4pytorch_b200 import _capi
    lib = _capi.bind(ctypes.CDLL(built_lib))
    d = _capi.make_desc(1, 1, 1, (229, 59), (12, 12), "gelu")
    n = lib.deno_spectral_twiddle_bytes(ctypes.byref(d))
    buf = np.zeros(n // 4, dtype=np.float32)
    assert lib.deno_spectral_twiddle_fill(ctypes.byref(d), buf.ctypes.data_as(ctypes.c_void_p), n) == 0
    last = buf[:59 * 12 * 2].reshape(59, 12, 2)
    ang = 2 * np.pi * np.outer(np.arange(59), np.arange(12)) / 59
    assert np.abs(last[..., 0] - np.cos(ang)).max() < 1e-7 and np.abs(last[..., 1] - np.sin(ang)).max() < 1e-7
    mid = buf[59 * 12 * 2:59 * 12 * 2 + 229 * 24 * 2].reshape(229, 24, 2)
    f = np.concatenate([np.arange(12), np.arange(229 - 12, 229)])
    ang = 2 * np.pi * np.outer(np.arange(229), f) / 229
    assert np.abs(mid[..., 0] - np.cos(ang)).max() < 1e-7 and np.abs(mid[..., 1] - np.sin(ang)).max() < 1e-7
    off = 59 * 12 * 2 + 229 * 24 * 2
    cl = buf[off:off + 12]
    assert cl.tolist() == [1.0] + [2.0] * 11
    # tensor-core twiddle images: hi + lo reproduces the fp64 value to ~1e-11, hi has a tf32 mantissa
    ae = buf[off + 12:off + 12 + 2 * 24 * 128]
    assert ae.size == 1 * 2 * 24 * 128
    hi, lo = ae[:24 * 128], ae[24 * 128:]
    assert (hi.view(np.uint32) & 0x1FFF).max() == 0
    y, l = 37, 5
    o = ((y & 7) * 16 + (y >> 3) * 128 + ((2 * l) & 3) * 4 + ((2 * l) >> 2) * 2048) // 4
    want = np.cos(2 * np.pi * l * y / 59)
    assert abs(float(hi[o]) + float(lo[o]) - want) < 1e-10
    # row_synthesis_tc A operand (last region of the table): two 128-row tiles of [hi | lo] images with KP = 48 columns,
    # column k < 24: cos(2 pi f_k x / H), column 24 + k: sin(...); rows beyond the grid are zero
    KP, ntm = 48, 2
    rs = buf[buf.size - ntm * 2 * KP * 128:]
    for x, k, part in ((5, 3, 0), (100, 17, 1), (128 + 77, 23, 1), (228, 0, 0), (128 + 101, 4, 0)):
        tm, r, kp = x // 128, x % 128, k + 24 * part
        o = ((r & 7) * 16 + (r >> 3) * 128 + (kp & 3) * 4 + (kp >> 2) * 2048) // 4
        hi_t = rs[tm * 2 * KP * 128:tm * 2 * KP * 128 + KP * 128]
        lo_t = rs[tm * 2 * KP * 128 + KP * 128:(tm + 1) * 2 * KP * 128]
        assert (hi_t.view(np.uint32) & 0x1FFF).max() == 0
        want = 0.0 if x >= 229 else (np.sin if part else np.cos)(2 * np.pi * f[k] * x / 229)
        assert abs(float(hi_t[o]) + float(lo_t[o]) - want) < 1e-10, (x, k, part)


@pytest.mark.parametrize("name", golden_names("model"))
def test_dropin_state_dict_matches_reference(name):
    import deno4pytorch_b200 as d4
    g = load_golden(name)
    m = dict(g["meta"])
    nd = m.pop("ndim")
    m.pop("kind")
    model = {1: d4.FNO1d, 2: d4.FNO2d, 3: d4.FNO3d}[nd](**m)
    sd = model.state_dict()
    assert list(sd.keys()) == list(g["params"].keys())
    for k, v in sd.items():
        assert v.shape == g["params"][k].shape and v.dtype == g["params"][k].dtype, k
    model.load_state_dict(g["params"], strict=True)


def test_layer_attributes_and_init_distribution():
    import deno4pytorch_b200 as d4
    torch.manual_seed(0)
    l2 = d4.SpectralConv2d(8, 6, (3, 5))
    assert (l2.modes1, l2.modes2, l2.in_dim, l2.out_dim, l2.norm, l2.return_freq, l2.use_complex) == \
        (3, 5, 8, 6, "ortho", False, True)
    assert isinstance(l2.dropout, torch.nn.Dropout) and l2.dropout.p == 0.1
    assert isinstance(l2.activation, torch.nn.GELU) and isinstance(l2.linear, torch.nn.Conv2d)
    w = torch.view_as_real(l2.weights1.detach())
    assert w.min() >= 0 and w.max() < l2.scale and abs(w.mean().item() - l2.scale / 2) < 0.1 * l2.scale
    assert d4.SpectralConv2d(4, 4, 7).modes2 == 7
    l1 = d4.SpectralConv1d(4, 4, 5, use_complex=False)
    assert l1.modes == 5 and l1.weights1.shape == (4, 4, 5, 2) and l1.weights1.dtype == torch.float32
    l3 = d4.SpectralConv3d(2, 3, (2, 3, 4), use_complex=False)
    assert isinstance(l3.activation, torch.nn.SiLU) and l3.weights4.shape == (2, 3, 2, 3, 4, 2)
    assert [k for k, _ in l3.named_parameters()] == ["weights1", "weights2", "weights3", "weights4", "linear.weight",
                                                     "linear.bias"]


def test_import_aliases_resolve_to_the_same_classes():
    code = (
        "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from fno.FNOs import FNO2d as A\n"
        "from Models.fno.FNOs import FNO2d as B\n"
        "import deno4pytorch_b200 as d4\n"
        "from Models.fno.spectral_layers import *\n"
        "assert A is B is d4.FNO2d, (A, B)\n"
        "assert nn and F and torch and activation_dict and SpectralConv3d\n"
        "import fno.spectral_layers as s; assert s.SpectralConv1d is d4.SpectralConv1d\n"
        "print('ok')\n" % (ROOT, os.path.join(ROOT, "Models")))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp")
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    from deno4pytorch_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.LibraryMissing, match="no CPU or PyTorch fallback"):
        _lib.lib()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "deno4pytorch_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "tests/emu" not in text or f.endswith((".cu", ".cuh")), f


_DP_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
from deno4pytorch_b200.parallel import setup_ddp, FlatGradients, shard_batch
rank, local_rank, world, device = setup_ddp(backend="gloo")
assert world == 2 and device.type == "cpu"
torch.manual_seed(0)
w_c = torch.nn.Parameter(torch.randn(3, 2, 4, dtype=torch.cfloat))
w_r = torch.nn.Parameter(torch.randn(5, 3))
flat = FlatGradients([w_c, w_r])
assert flat.flat.numel() == 24 * 2 + 16
gx = torch.randn(8, 3, generator=torch.Generator().manual_seed(1))
xs = shard_batch(gx, rank, world)
flat.zero()
loss = (xs @ w_r.t()).pow(2).sum() + (torch.view_as_real(w_c) * (rank + 1)).sum()
loss.backward()
assert w_r.grad.data_ptr() >= flat.flat.data_ptr()          # autograd accumulated into the views
local = flat.flat.clone()
flat.all_reduce_mean()
gathered = [torch.zeros_like(local) for _ in range(world)]
dist.all_gather(gathered, local)
want = sum(gathered) / world
assert torch.allclose(flat.flat, want, rtol=1e-6, atol=1e-6)
assert torch.allclose(torch.view_as_real(w_c.grad), torch.full((3, 2, 4, 2), 1.5))
dist.barrier()
if rank == 0:
    print("dp-ok")
"""


def test_flat_gradient_allreduce_two_ranks_gloo(tmp_path):
    script = tmp_path / "dp_worker.py"
    script.write_text(_DP_WORKER.format(root=ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29613", str(script)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0 and "dp-ok" in out.stdout, out.stdout + out.stderr


def test_flat_gradient_spans_and_direct_write_registration():
    """Host-side bookkeeping of parallel.FlatGradients (no kernels involved): per-stage slices of the flat buffer and
    the gradient-sink registration the backward kernels look up (functional.grad_sink)."""
    import torch
    from deno4pytorch_b200.functional import grad_sink
    from deno4pytorch_b200.parallel import FlatGradients
    a = torch.nn.Parameter(torch.zeros(3, 5))
    b = torch.nn.Parameter(torch.zeros(7, dtype=torch.complex64))
    c = torch.nn.Parameter(torch.zeros(2))
    fg = FlatGradients([a, b, c])
    assert fg.flat.numel() == 16 + 16 + 4                      # every span padded to a multiple of 4 floats
    assert fg.span_of([a, b]).numel() == 32 and fg.span_of([c]).numel() == 4
    assert fg.span_of([a, b]).data_ptr() == fg.flat.data_ptr()
    with pytest.raises(ValueError):
        fg.span_of([a, c])                                     # not adjacent
    assert grad_sink(a) is None
    fg.direct_writes(True)
    assert grad_sink(a) is a.grad and grad_sink(b) is b.grad
    assert grad_sink(a, like=torch.zeros(5, 3)) is None         # shape mismatch: no sink
    b.grad.real.fill_(2.0)
    assert float(fg.flat[16:30:2].sum()) == 14.0                # complex gradients live in the flat buffer as (re, im)
    fg.direct_writes(False)
    assert grad_sink(a) is None


# ---- §8f rank 4: the adaptive Fourier layers and the 4-D stand-alone model as drop-ins (host side only) --------------------
@pytest.mark.parametrize("name", golden_names("afno"))
def test_afno_dropin_state_dict_and_attributes_match_reference(name):
    import deno4pytorch_b200 as d4
    g = load_golden(name)
    m = g["meta"]
    cls = {1: d4.AdaptiveFourier1d, 2: d4.AdaptiveFourier2d, 3: d4.AdaptiveFourier3d}[m["ndim"]]
    layer = cls(m["hidden_size"], num_blocks=m["num_blocks"], sparsity_threshold=m["sparsity_threshold"],
                hard_thresholding_fraction=m["hard_thresholding_fraction"], activation=m["activation"])
    sd = layer.state_dict()
    assert list(sd.keys()) == list(g["params"].keys()) == ["weights1", "weights2", "bias1", "bias2"]
    for k, v in sd.items():
        assert v.shape == g["params"][k].shape and v.dtype == g["params"][k].dtype == torch.complex64, k
    layer.load_state_dict(g["params"], strict=True)
    assert (layer.hidden_size, layer.num_blocks, layer.block_size, layer.hidden_size_factor, layer.scale) == \
        (m["hidden_size"], m["num_blocks"], m["hidden_size"] // m["num_blocks"], 1, 0.02)


def test_afno_defaults_and_init_distribution():
    import deno4pytorch_b200 as d4
    torch.manual_seed(0)
    l1, l2, l3 = d4.AdaptiveFourier1d(64), d4.AdaptiveFourier2d(64), d4.AdaptiveFourier3d(64, num_blocks=4)
    assert isinstance(l1.activation, torch.nn.ReLU) and isinstance(l2.activation, torch.nn.GELU) \
        and isinstance(l3.activation, torch.nn.GELU)                      # spectral_layers.py:350, :426, :504
    assert (l2.num_blocks, l2.sparsity_threshold, l2.hard_thresholding_fraction) == (8, 0.01, 1)
    assert l3.weights1.shape == (4, 16, 16) and l3.bias2.shape == (4, 16)
    w = torch.view_as_real(l2.weights1.detach())
    assert w.min() >= 0 and w.max() < 0.02 and abs(w.mean().item() - 0.01) < 1e-3      # 0.02 * U[0, 1) on re and im
    with pytest.raises(AssertionError, match="divisble"):
        d4.AdaptiveFourier2d(10, num_blocks=4)
    from deno4pytorch_b200.functional import afno_kept_modes
    assert afno_kept_modes((55, 64), 1) == 33 and afno_kept_modes((16,), 0.5) == 4 and afno_kept_modes((4, 4), 0.01) == 0


def test_fno4d_dropin_state_dict_matches_reference():
    from deno4pytorch_b200.FNO_HIT import FNO4d, SpectralConv4d
    g = load_golden("fno4d")
    m = g["meta"]
    model = FNO4d(*m["modes"][:3], 4, m["width"])
    assert list(model.state_dict().keys()) == list(g["params"].keys())
    model.load_state_dict(g["params"], strict=True)
    gl = load_golden("layer4d")
    layer = SpectralConv4d(gl["meta"]["in_dim"], gl["meta"]["out_dim"], *gl["meta"]["modes"])
    layer.load_state_dict(gl["params"], strict=True)
    assert layer.modes4 == 2 and SpectralConv4d(2, 2, 4, 4, 4, 9).modes4 == 2      # clamped like FNO_HIT.py:44
