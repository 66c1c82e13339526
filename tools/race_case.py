"""One forward + backward of the (128, 40) / modes (64, 16) layer on the SIMT kernels - small enough for
compute-sanitizer --tool racecheck / memcheck / initcheck.  python tools/race_case.py [path]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from deno4pytorch_b200.functional import spectral_conv
os.environ["DENO_SPECTRAL_PATH"] = sys.argv[1] if len(sys.argv) > 1 else "simt"; __import__("deno4pytorch_b200.functional", fromlist=["reload_env"]).reload_env()
B, cin, cout, spatial, modes, act = (2, 6, 6, (128, 40), (64, 16), "tanh")
gen = torch.Generator().manual_seed(11)
dev = torch.device("cuda", 0)
x = torch.randn((B, cin) + spatial, generator=gen).to(dev).requires_grad_(True)
ws = [torch.view_as_complex(torch.randn((cin, cout) + modes + (2,), generator=gen) / cin).to(dev).requires_grad_(True) for _ in range(2)]
lw = (torch.randn((cout, cin, 1, 1), generator=gen) / cin ** 0.5).to(dev).requires_grad_(True)
lb = torch.randn(cout, generator=gen).to(dev).requires_grad_(True)
cot = torch.randn((B, cout) + spatial, generator=gen).to(dev)
y = spectral_conv(x, ws, lw, lb, modes, act)
(y * cot).sum().backward()
torch.cuda.synchronize()
print("done", float(y.abs().sum()), float(x.grad.abs().sum()))
