import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from golden_util import rel_l2
from oracle import fno_port
from deno4pytorch_b200.functional import spectral_conv
B, cin, cout, spatial, modes, act = (2, 6, 6, (128, 40), (64, 16), "tanh")
nd = 2
gen = torch.Generator().manual_seed(11)
x = torch.randn((B, cin) + spatial, generator=gen)
ws = [torch.view_as_complex(torch.randn((cin, cout) + modes + (2,), generator=gen) / cin) for _ in range(2)]
lw = torch.randn((cout, cin) + (1,) * nd, generator=gen) / cin ** 0.5
lb = torch.randn(cout, generator=gen)
cot = torch.randn((B, cout) + spatial, generator=gen)
xr = x.clone().requires_grad_(True); wr = [w.clone().requires_grad_(True) for w in ws]
lwr, lbr = lw.clone().requires_grad_(True), lb.clone().requires_grad_(True)
ref = fno_port.spectral_conv(xr, wr, lwr, lbr, modes, activation=act); (ref * cot).sum().backward()
x64 = x.double().requires_grad_(True); w64 = [w.to(torch.complex128).requires_grad_(True) for w in ws]
r64 = fno_port.spectral_conv(x64, w64, lw.double(), lb.double(), modes, activation=act); (r64 * cot.double()).sum().backward()
print("fp32 oracle vs fp64 oracle: y %.2e gx %.2e" % (rel_l2(ref, r64), rel_l2(xr.grad, x64.grad)))
dev = torch.device("cuda", 0)
for path in ["simt", "auto", "simt", "auto"]:
    os.environ["DENO_SPECTRAL_PATH"] = path; __import__("deno4pytorch_b200.functional", fromlist=["reload_env"]).reload_env()
    xg = x.to(dev).requires_grad_(True); wg = [w.to(dev).requires_grad_(True) for w in ws]
    lwg = lw.to(dev).requires_grad_(True); lbg = lb.to(dev).requires_grad_(True)
    y = spectral_conv(xg, wg, lwg, lbg, modes, act); (y * cot.to(dev)).sum().backward(); torch.cuda.synchronize()
    print(path, "vs fp32 oracle: y %.2e gx %.2e gw %.2e | vs fp64: y %.2e gx %.2e" % (
        rel_l2(y.cpu(), ref), rel_l2(xg.grad.cpu(), xr.grad), rel_l2(wg[0].grad.cpu(), wr[0].grad),
        rel_l2(y.cpu(), r64), rel_l2(xg.grad.cpu(), x64.grad)))

# nondeterminism hunt: the GPU path repeated (bitwise) and the fp32 CPU oracle repeated
os.environ["DENO_SPECTRAL_PATH"] = "simt"; __import__("deno4pytorch_b200.functional", fromlist=["reload_env"]).reload_env()
first = None
ndiff = 0
worst = 0.0
for it in range(60):
    xg = x.to(dev).requires_grad_(True); wg = [w.to(dev).requires_grad_(True) for w in ws]
    lwg = lw.to(dev).requires_grad_(True); lbg = lb.to(dev).requires_grad_(True)
    y = spectral_conv(xg, wg, lwg, lbg, modes, act); (y * cot.to(dev)).sum().backward(); torch.cuda.synchronize()
    cur = (y.detach().clone(), xg.grad.clone())
    worst = max(worst, rel_l2(cur[1].cpu(), x64.grad))
    if first is None:
        first = cur
    elif not (torch.equal(first[0], cur[0]) and torch.equal(first[1], cur[1])):
        ndiff += 1
print("GPU simt path, 60 repetitions: %d differ bitwise from the first; worst gx vs fp64 %.2e" % (ndiff, worst))
errs = []
for it in range(12):
    xr2 = x.clone().requires_grad_(True)
    r2 = fno_port.spectral_conv(xr2, [w.clone() for w in ws], lw.clone(), lb.clone(), modes, activation=act)
    (r2 * cot).sum().backward()
    errs.append(rel_l2(xr2.grad, x64.grad))
print("fp32 CPU oracle gx vs fp64, 12 repetitions: min %.2e max %.2e" % (min(errs), max(errs)))
