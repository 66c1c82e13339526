"""Bring-up of plane_analysis_tma_kernel: one forward (and optionally backward) layer call per sub-process with
DENO_ANALYSIS_TMA=1 and DENO_AT_DBG=<bits> (1 no L2 prefetch, 2 no MMA issue, 4 no TMA loads, 8 no TMEM loads), so
an illegal instruction kills only that sub-process.  python tools/an_tma_debug.py [child <dbg> <B> <C> <H> <W> <m> <bwd>]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(dbg, B, C, H, W, m, bwd):
    import torch
    from deno4pytorch_b200.functional import spectral_conv, reload_env
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    x = torch.randn(B, C, H, W, device=dev, requires_grad=True)
    ws = [(torch.rand(C, C, m, m, dtype=torch.cfloat, device=dev) / (C * C)).requires_grad_(True) for _ in range(2)]
    lw = (torch.randn(C, C, 1, 1, device=dev) / C ** 0.5).requires_grad_(True)
    lb = torch.randn(C, device=dev, requires_grad=True)
    res = {}
    for flag in ("0", "1"):
        os.environ["DENO_ANALYSIS_TMA"] = flag
        os.environ["DENO_AT_DBG"] = str(dbg)
        reload_env()
        y = spectral_conv(x, ws, lw, lb, (m, m), "gelu")
        if bwd:
            g = torch.autograd.grad(y.square().sum(), [x, lw] + ws)
            res[flag] = [y.detach()] + [t.detach() for t in g]
        else:
            res[flag] = [y.detach()]
        torch.cuda.synchronize()
    errs = [float((a - b).abs().max() / (b.abs().max() + 1e-30)) for a, b in zip(res["1"], res["0"])]
    print("rel max err vs default kernels:", ["%.2e" % e for e in errs], flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child(*[int(a) for a in sys.argv[2:9]])
        sys.exit(0)
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    shapes = [(2, 8, 16, 16, 4), (2, 32, 94, 94, 12), (3, 16, 72, 72, 12), (2, 8, 20, 63, 5), (20, 32, 94, 94, 12)]
    for (B, C, H, W, m) in shapes:
        for bwd in (0, 1):
            for dbg in ((0,) if quick else (0, 1, 3, 5, 9, 15, 7, 11, 13)):
                try:
                    r = subprocess.run([sys.executable, __file__, "child", str(dbg), str(B), str(C), str(H), str(W), str(m), str(bwd)],
                                       capture_output=True, text=True, timeout=60)
                except subprocess.TimeoutExpired:
                    print(f"shape {(B, C, H, W, m)} bwd={bwd} dbg={dbg:2d}: TIMEOUT (hang)", flush=True)
                    continue
                tail = (r.stdout.strip().splitlines() or [""])[-1]
                err = [l for l in r.stderr.splitlines() if "error" in l.lower()][:1]
                print(f"shape {(B, C, H, W, m)} bwd={bwd} dbg={dbg:2d}: rc={r.returncode} {tail} {err}", flush=True)
