"""Which batch entries reach the TMA bypass weight gradient?  x[b] = 10^b, gy = 1 -> gW[o, i] / S = sum_b 10^b over the
entries that contributed (decimal digits = presence flags).  python tools/wgrad_diag2.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from deno4pytorch_b200.functional import spectral_conv, reload_env

dev = torch.device("cuda", 0)
for C, sp in ((32, (8, 8)), (32, (16, 16)), (16, (8, 12)), (64, (8, 8))):
    for B in (1, 2, 3, 4, 5, 6, 7, 8):
        S = sp[0] * sp[1]
        x = (10.0 ** torch.arange(B, dtype=torch.float64)).float().view(B, 1, 1, 1).expand(B, C, *sp).contiguous()
        gy = torch.ones(B, C, *sp)
        ws = [torch.zeros(C, C, 2, 2, dtype=torch.cfloat, device=dev, requires_grad=True) for _ in range(2)]
        out = {}
        for flag in ("0", "1"):
            os.environ["DENO_WGRAD_TMA"] = flag
            os.environ["DENO_ANALYSIS_TMA"] = "0"
            reload_env()
            lw = torch.zeros(C, C, 1, 1, device=dev, requires_grad=True)
            xg = x.to(dev).requires_grad_(True)
            y = spectral_conv(xg, ws, lw, None, (2, 2), "identity")
            y.backward(gy.to(dev))
            torch.cuda.synchronize()
            g = lw.grad.view(C, C).cpu().double() / S
            vals = sorted(set(int(round(v)) for v in g.flatten().tolist()))
            out[flag] = vals
        print(f"C={C} S={S} B={B}: old {out['0']}  tma {out['1']}", flush=True)
