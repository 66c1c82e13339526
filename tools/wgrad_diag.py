"""Structured-input check of the bypass weight gradient (TMA kernel vs the register-staged kernel vs the exact answer):
gy one-hot in (o, pt), x = i + 100 * pt + 10000 * b, identity activation.  python tools/wgrad_diag.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from deno4pytorch_b200.functional import spectral_conv, reload_env

dev = torch.device("cuda", 0)
torch.set_printoptions(linewidth=250, precision=1, sci_mode=False)
CASES = [(4, 32, (8, 8)), (4, 32, (8, 16)), (8, 32, (8, 8)), (4, 32, (64, 100)), (4, 32, (128, 100)), (3, 32, (8, 8)), (1, 32, (8, 8)),
         (2, 64, (8, 8)), (4, 64, (64, 100)), (20, 32, (94, 94))]
for B, C, sp in CASES:
    S = sp[0] * sp[1]
    x = (torch.arange(C).view(1, C, 1) + 100.0 * (torch.arange(S) % 64).view(1, 1, S) + 10000.0 * torch.arange(B).view(B, 1, 1)).view(B, C, *sp)
    gy = torch.zeros(B, C, S)
    for b in range(B):
        for o in range(C):
            gy[b, o, (o + b) % S] = 1.0
            gy[b, o, (S - 1 - o - b) % S] += 0.5
    gy = gy.view(B, C, *sp)
    ws = [torch.zeros(C, C, 2, 2, dtype=torch.cfloat, device=dev, requires_grad=True) for _ in range(2)]
    res = {}
    for flag in ("0", "1"):
        os.environ["DENO_WGRAD_TMA"] = flag
        os.environ["DENO_ANALYSIS_TMA"] = "0"
        reload_env()
        lw = torch.zeros(C, C, 1, 1, device=dev, requires_grad=True)
        xg = x.to(dev).requires_grad_(True)
        y = spectral_conv(xg, ws, lw, None, (2, 2), "identity")
        y.backward(gy.to(dev))
        torch.cuda.synchronize()
        res[flag] = lw.grad.view(C, C).cpu().double()
    exp = torch.einsum("bop,bip->oi", gy.view(B, C, S).double(), x.view(B, C, S).double())
    e0 = float((res['0'] - exp).abs().max() / exp.abs().max())
    e1 = float((res['1'] - exp).abs().max() / exp.abs().max())
    nz = int((res['1'] != 0).sum())
    print(f"B={B} C={C} S={S} chunks={-(-S // 32)}: old rel err {e0:.2e}, TMA rel err {e1:.2e}, TMA nonzeros {nz}/{C * C}, "
          f"sum ratio {float(res['1'].sum() / exp.sum()):.4f}")
    if e1 > 1e-4 and C * C <= 4096 and nz > 0:
        r = res["1"] / exp
        print("  ratio rows 0,1,8,31 cols 0..7:", [[round(float(v), 3) for v in r[i, :8]] for i in (0, 1, 8, 31)])
