"""fp64 truncated-DFT statement of the spectral layer, forward and explicit adjoints
(TEST INFRASTRUCTURE ONLY; numpy, no FFT calls).

This is the "math contract" of SURVEY.md section 8(a): the reference computes a full
``rfftn``, keeps the corner modes, mixes channels per mode, zero-pads and runs a full
``irfftn`` (/root/reference/Models/fno/spectral_layers.py:90-110,189-203,305-333).
Algebraically that is a *truncated* forward DFT onto the retained modes, the per-mode
``Cin x Cout`` complex product, and a truncated inverse with Hermitian weights
``c_l`` on the half-spectrum axis (``irfft`` discards the imaginary part of the
``l = 0`` and Nyquist bins and counts every other retained ``l`` twice).  The CUDA
kernels implement exactly these sums, so every kernel stage can be pinned against
the intermediate arrays returned here.

Parity status: PINNED through ``tests/test_oracle_golden.py`` (agreement with
``oracle/fno_port.py`` and with the reference fixtures, forward and every gradient).

Packed-spectrum layout used everywhere in this repo: for a leading axis of length
``N`` with ``m`` retained modes the packed index ``p in [0, 2m)`` maps to frequency
``p`` for ``p < m`` and ``N - 2m + p`` otherwise (i.e. the negative frequencies
``-m..-1``); the last axis keeps ``l in [0, m)``.  Weight blocks ``weights1..4`` are
concatenated along the leading packed axes in the reference's block order
(spectral_layers.py:196-199, 318-325).
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np
from scipy.special import erf

_AX = "xyz"
_PK = "uvw"


def packed_freqs(n: int, m: int, half: bool) -> np.ndarray:
    if half:
        assert m <= n // 2 + 1, "more retained modes than rfft bins"
        return np.arange(m)
    assert 2 * m <= n, "overlapping low/high mode blocks are not covered by the truth model"
    return np.concatenate([np.arange(m), np.arange(n - m, n)])


def hermitian_weights(n_last: int, m_last: int) -> np.ndarray:
    """c_l of the C2R transform: 1 at l=0 (and at Nyquist when n_last is even), else 2."""
    c = np.full(m_last, 2.0)
    c[0] = 1.0
    if n_last % 2 == 0 and m_last - 1 == n_last // 2:
        c[-1] = 1.0
    return c


def dft_matrices(spatial: Sequence[int], modes: Sequence[int]) -> List[np.ndarray]:
    """E_a[n, p] = exp(-2 pi i f_p n / N_a) for every axis a."""
    mats = []
    nd = len(spatial)
    for a, (n, m) in enumerate(zip(spatial, modes)):
        f = packed_freqs(n, m, half=(a == nd - 1))
        ang = (np.outer(np.arange(n), f) % n) * (2.0 * np.pi / n)
        mats.append(np.exp(-1j * ang))
    return mats


def pack_weights(weights: Sequence[np.ndarray]) -> np.ndarray:
    """Blocks (Cin,Cout,*modes) complex -> (Cin,Cout,2m1[,2m2],m_last)."""
    ws = [np.asarray(w) for w in weights]
    ws = [w[..., 0] + 1j * w[..., 1] if not np.iscomplexobj(w) else w for w in ws]
    nd = ws[0].ndim - 2
    assert len(ws) == 2 ** (nd - 1)
    if nd == 1:
        return ws[0].astype(np.complex128)
    if nd == 2:
        return np.concatenate([ws[0], ws[1]], axis=2).astype(np.complex128)
    lo_y = np.concatenate([ws[0], ws[1]], axis=2)   # (+x,+y), (-x,+y)
    hi_y = np.concatenate([ws[2], ws[3]], axis=2)   # (+x,-y), (-x,-y)
    return np.concatenate([lo_y, hi_y], axis=3).astype(np.complex128)


def unpack_weight_grad(gwc: np.ndarray, modes: Sequence[int]) -> List[np.ndarray]:
    nd = len(modes)
    if nd == 1:
        return [gwc]
    m1 = modes[0]
    if nd == 2:
        return [gwc[:, :, :m1], gwc[:, :, m1:]]
    m2 = modes[1]
    return [gwc[:, :, :m1, :m2], gwc[:, :, m1:, :m2], gwc[:, :, :m1, m2:], gwc[:, :, m1:, m2:]]


def act_and_grad(name, z: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Activation table of /root/reference/Models/configs.py:17-19 with derivatives."""
    if name == "gelu":
        cdf = 0.5 * (1.0 + erf(z / np.sqrt(2.0)))
        pdf = np.exp(-0.5 * z * z) / np.sqrt(2.0 * np.pi)
        return z * cdf, cdf + z * pdf
    if name == "silu" or name is None:
        s = 1.0 / (1.0 + np.exp(-z))
        return z * s, s * (1.0 + z * (1.0 - s))
    if name == "relu":
        return np.maximum(z, 0.0), (z > 0).astype(z.dtype)
    if name == "tanh":
        t = np.tanh(z)
        return t, 1.0 - t * t
    if name == "leakyrelu":
        return np.where(z > 0, z, 0.01 * z), np.where(z > 0, 1.0, 0.01)
    if name == "identity":
        return z, np.ones_like(z)
    raise KeyError(name)


def _analysis(u: np.ndarray, mats: Sequence[np.ndarray]) -> np.ndarray:
    """(B,C,*spatial) real -> (B,C,*packed) complex: sum_n u * prod_a E_a[n_a,p_a]."""
    nd = len(mats)
    expr = "bc" + _AX[:nd] + "," + ",".join(_AX[a] + _PK[a] for a in range(nd)) + "->bc" + _PK[:nd]
    return np.einsum(expr, u.astype(np.complex128), *mats, optimize=True)


def _synthesis_real(spec: np.ndarray, mats: Sequence[np.ndarray]) -> np.ndarray:
    """(B,C,*packed) complex -> Re sum_p spec * prod_a conj(E_a[n_a,p_a]), (B,C,*spatial)."""
    nd = len(mats)
    expr = "bc" + _PK[:nd] + "," + ",".join(_AX[a] + _PK[a] for a in range(nd)) + "->bc" + _AX[:nd]
    return np.einsum(expr, spec, *[np.conj(m) for m in mats], optimize=True).real


def forward(x, weights, lin_w, lin_b, modes, activation="gelu") -> Dict[str, np.ndarray]:
    """Forward pass; returns every intermediate the kernels materialise."""
    x = np.asarray(x, dtype=np.float64)
    nd = x.ndim - 2
    spatial = x.shape[2:]
    modes = (modes,) * nd if isinstance(modes, int) else tuple(modes)
    if nd == 1:
        assert spatial[0] % 2 == 0, "reference SpectralConv1d fails for odd length (irfft without n=)"
    mats = dft_matrices(spatial, modes)
    wc = pack_weights(weights)
    c_l = hermitian_weights(spatial[-1], modes[-1])
    s_total = float(np.prod(spatial))

    xhat = _analysis(x, mats)                                   # X^[b,i,p]
    yhat = np.einsum("bi...,io...->bo...", xhat, wc)            # Y^[b,o,p]
    spec = _synthesis_real(yhat * c_l, mats) / s_total          # s[b,o,n]
    lw = np.asarray(lin_w, dtype=np.float64).reshape(lin_w.shape[0], lin_w.shape[1])
    conv = np.einsum("oi,bi...->bo...", lw, x)
    if lin_b is not None:
        conv = conv + np.asarray(lin_b, dtype=np.float64).reshape((1, -1) + (1,) * nd)
    z = spec + conv
    y, dact = act_and_grad(activation, z)
    return dict(xhat=xhat, yhat=yhat, spec=spec, z=z, y=y, dact=dact, mats=mats, wc=wc, c_l=c_l,
                s_total=s_total, lw=lw, modes=modes)


def backward(x, fwd: Dict[str, np.ndarray], gy) -> Dict[str, np.ndarray]:
    """Explicit adjoints.  Weight gradients follow PyTorch's convention
    for complex leaves: ``.grad = dL/dRe + i dL/dIm``."""
    x = np.asarray(x, dtype=np.float64)
    gy = np.asarray(gy, dtype=np.float64)
    nd = x.ndim - 2
    mats, wc, c_l = fwd["mats"], fwd["wc"], fwd["c_l"]
    gz = gy * fwd["dact"]
    ghat = _analysis(gz, mats) * (c_l / fwd["s_total"])                      # G^[b,o,p]
    gwc = np.einsum("bi...,bo...->io...", np.conj(fwd["xhat"]), ghat)         # gWc[i,o,p]
    gxhat = np.einsum("io...,bo...->bi...", np.conj(wc), ghat)                # G^x[b,i,p]
    gx = _synthesis_real(gxhat, mats) + np.einsum("oi,bo...->bi...", fwd["lw"], gz)
    red = (0,) + tuple(range(2, 2 + nd))
    glw = np.einsum("bos,bis->oi", gz.reshape(gz.shape[0], gz.shape[1], -1), x.reshape(x.shape[0], x.shape[1], -1))
    glb = gz.sum(axis=red)
    return dict(gz=gz, ghat=ghat, gxhat=gxhat, gx=gx, gwc=gwc,
                gweights=unpack_weight_grad(gwc, fwd["modes"]), glin_w=glw, glin_b=glb)
