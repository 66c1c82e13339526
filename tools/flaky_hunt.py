"""Repeats the config-shape parity cases of tests/test_gpu_parity.py in one process (same allocator reuse pattern as the
suite) and reports every deviation of a GPU result from its own first run, plus the error against the CPU oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from golden_util import rel_l2
from oracle import fno_port
from deno4pytorch_b200.functional import spectral_conv

SHAPES = [
    (3, 64, 64, (1026,), (16,), "gelu"), (2, 32, 32, (94, 94), (12, 12), "gelu"), (2, 64, 64, (72, 72), (12, 12), "gelu"),
    (1, 32, 32, (70, 70, 46), (8, 8, 8), "gelu"), (2, 5, 7, (33, 20), (16, 11), "silu"), (2, 6, 6, (128, 40), (64, 16), "tanh"),
]
dev = torch.device("cuda", 0)
cases = []
for shape in SHAPES:
    B, cin, cout, spatial, modes, act = shape
    nd = len(spatial)
    gen = torch.Generator().manual_seed(11)
    x = torch.randn((B, cin) + spatial, generator=gen)
    ws = [torch.view_as_complex(torch.randn((cin, cout) + modes + (2,), generator=gen) / cin) for _ in range(2 ** (nd - 1))]
    lw = torch.randn((cout, cin) + (1,) * nd, generator=gen) / cin ** 0.5
    lb = torch.randn(cout, generator=gen)
    cot = torch.randn((B, cout) + spatial, generator=gen)
    xr = x.clone().requires_grad_(True)
    ref = fno_port.spectral_conv(xr, [w.clone() for w in ws], lw, lb, modes, activation=act)
    (ref * cot).sum().backward()
    cases.append((shape, x, ws, lw, lb, cot, ref.detach(), xr.grad.clone()))
first = {}
bad = 0
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 25
for rep in range(reps):
    for path in ("simt", "auto"):
        os.environ["DENO_SPECTRAL_PATH"] = path; __import__("deno4pytorch_b200.functional", fromlist=["reload_env"]).reload_env()
        for ci, (shape, x, ws, lw, lb, cot, ref, gref) in enumerate(cases):
            xg = x.to(dev).requires_grad_(True)
            wg = [w.to(dev).requires_grad_(True) for w in ws]
            lwg = lw.to(dev).requires_grad_(True); lbg = lb.to(dev).requires_grad_(True)
            y = spectral_conv(xg, wg, lwg, lbg, shape[4], shape[5])
            (y * cot.to(dev)).sum().backward()
            torch.cuda.synchronize()
            cur = [y.detach(), xg.grad, lwg.grad, lbg.grad] + [torch.view_as_real(w.grad) for w in wg]
            key = (path, ci)
            if key not in first:
                first[key] = [t.clone() for t in cur]
                print(path, shape[3], "y %.2e gx %.2e" % (rel_l2(y.cpu(), ref), rel_l2(xg.grad.cpu(), gref)), flush=True)
            else:
                for name, a, b in zip(["y", "gx", "glw", "glb", "gw0", "gw1", "gw2", "gw3"], cur, first[key]):
                    if not torch.equal(a, b):
                        bad += 1
                        d = (a - b).abs()
                        print("rep %d %s %s: %s differs from first run: rel %.3e, %d elements, max |d| %.3e at %s" % (
                            rep, path, shape[3], name, rel_l2(a.cpu(), b.cpu()), int((d > 0).sum()), float(d.max()),
                            tuple(int(v) for v in torch.nonzero(d == d.max())[0])), flush=True)
print("repetitions", reps, "deviations", bad)
