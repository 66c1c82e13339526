"""Row n3 of the coverage table (north_star: "the choice between partial-DFT-as-GEMM and a radix FFT is made per shape
from measured counters"): the full-length radix FFT the reference calls (cuFFT through torch.fft.rfftn / irfftn,
/root/reference/Models/fno/spectral_layers.py:95,191,310 and :106,202,329) timed next to this library's truncated-DFT
analysis / synthesis kernels on the same tensors, same GPU, CUDA events, warm, 20 repetitions.

For every shape: cuFFT forward (rfftn of the whole tensor + the corner-mode gather the reference does), cuFFT inverse
(zero-padded irfftn of the full half-spectrum), and the device time of this library's analysis (forward transform
only) and of its row_synthesis + plane_synthesis launches (inverse transform + bypass conv + bias + GELU + act'(z), i.e.
MORE work than the irfftn alone).  Also the algorithmic byte ratio: a full half-spectrum is S/2 + 1 complex values per
channel (= A bytes again), the retained modes are M << S.

    python tools/fft_vs_dft.py > profiles/r2_fft_vs_dft.json
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

SHAPES = [
    ("C2 Darcy 94x94 (2*47, not FFT-friendly)", 20, 32, (94, 94), (12, 12)),
    ("C3 NS 72x72 (2^3 3^2, FFT-friendly)", 20, 64, (72, 72), (12, 12)),
    ("NS 64x64 padding 0 (2^6)", 20, 64, (64, 64), (12, 12)),
    ("C4b Pipe 137x137 (prime)", 20, 128, (137, 137), (12, 12)),
    ("C5 turbulence 70x70x46", 10, 32, (70, 70, 46), (8, 8, 8)),
    ("3-D 64^3 padding 0 (2^6 per axis)", 4, 32, (64, 64, 64), (8, 8, 8)),
    ("C1 Burgers 1026", 20, 64, (1026,), (16,)),
]


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3   # us


def main():
    import ctypes as C
    from deno4pytorch_b200 import _capi
    from deno4pytorch_b200._lib import lib
    from deno4pytorch_b200.functional import spectral_conv
    dev = torch.device("cuda", 0)
    L = lib()
    out = {"what": __doc__.strip().splitlines()[0], "gpu": torch.cuda.get_device_name(0), "shapes": []}
    for name, B, Cw, sp, modes in SHAPES:
        nd = len(sp)
        dims = tuple(range(-nd, 0))
        x = torch.randn((B, Cw) + sp, device=dev)
        S = 1
        for s in sp:
            S *= s
        nblk = 2 ** (nd - 1)
        M = nblk
        for m in modes:
            M *= m
        A = 4 * B * Cw * S
        spec_bytes = 8 * B * Cw * M

        def fft_fwd():
            xf = torch.fft.rfftn(x, dim=dims)
            if nd == 1:
                return xf[..., :modes[0]].contiguous()
            if nd == 2:
                return torch.cat((xf[:, :, :modes[0], :modes[1]], xf[:, :, -modes[0]:, :modes[1]]), dim=2)
            a = torch.cat((xf[:, :, :modes[0], :modes[1], :modes[2]], xf[:, :, -modes[0]:, :modes[1], :modes[2]]), dim=2)
            b = torch.cat((xf[:, :, :modes[0], -modes[1]:, :modes[2]], xf[:, :, -modes[0]:, -modes[1]:, :modes[2]]), dim=2)
            return torch.cat((a, b), dim=3)

        half = sp[:-1] + (sp[-1] // 2 + 1,)
        full_spec = torch.zeros((B, Cw) + half, dtype=torch.cfloat, device=dev)

        def fft_inv():
            return torch.fft.irfftn(full_spec, s=sp, dim=dims)

        t_fft_fwd = timed(fft_fwd)
        t_fft_only = timed(lambda: torch.fft.rfftn(x, dim=dims))
        t_fft_inv = timed(fft_inv)
        # this library: one forward layer call, per-kernel device times from the library's own event records
        ws = [torch.view_as_complex(torch.rand((Cw, Cw) + modes + (2,), device=dev) / (Cw * Cw)) for _ in range(nblk)]
        lw = torch.randn((Cw, Cw) + (1,) * nd, device=dev) / Cw ** 0.5
        lb = torch.randn(Cw, device=dev)
        xg = x.clone().requires_grad_(True)       # grad mode on: the forward also writes act'(z) (the training forward)
        for _ in range(3):
            spectral_conv(xg, ws, lw, lb, modes, "gelu")
        acc = {}
        recs = (_capi.ProfileRecord * 64)()
        reps = 20
        for _ in range(reps):
            torch.cuda.synchronize()
            L.deno_profile_begin()
            spectral_conv(xg, ws, lw, lb, modes, "gelu")
            torch.cuda.synchronize()
            n = L.deno_profile_end(recs, 64)
            for i in range(min(n, 64)):
                acc.setdefault(recs[i].name.decode(), []).append(recs[i].ms)
        k_us = {k: round(sum(v) / reps * 1e3, 2) for k, v in acc.items()}
        analysis = sum(v for k, v in k_us.items() if "analysis" in k)
        synthesis = sum(v for k, v in k_us.items() if "synthesis" in k)
        out["shapes"].append({
            "shape": name, "tensor": [B, Cw] + list(sp), "modes": list(modes), "A_bytes": A,
            "retained_spectrum_bytes": spec_bytes, "full_half_spectrum_bytes": 8 * B * Cw * (S // sp[-1]) * (sp[-1] // 2 + 1),
            "cufft_rfftn_us": round(t_fft_only, 2), "cufft_rfftn_plus_corner_gather_us": round(t_fft_fwd, 2),
            "cufft_irfftn_us": round(t_fft_inv, 2),
            "truncated_dft_analysis_us": round(analysis, 2),
            "truncated_dft_synthesis_incl_bypass_conv_bias_gelu_us": round(synthesis, 2),
            "library_forward_kernels_us": k_us,
            "forward_transform_speedup_vs_cufft": round(t_fft_fwd / analysis, 2) if analysis > 0 else None,
            "choice": "truncated DFT as GEMM" if analysis < t_fft_fwd else "radix FFT would win the forward transform",
        })
        del x, full_spec, xg
        torch.cuda.empty_cache()
    out["note"] = ("Per-launch event times carry ~5 us of event overhead each; cuFFT times are whole-op event times. The FFT "
                   "computes S/2+1 bins per line of which M are kept: its output alone is another activation-sized tensor "
                   "(full_half_spectrum_bytes ~ A) that the truncated DFT never writes.")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
