"""Full-batch parity at the EXACT shapes bench.py runs (BASELINE.json configs): forward and every
gradient of one spectral layer at C1 B20, C2 B20, C3 B20, C4b B20 and C5 B10, on both kernel families, against the
reference algorithm (oracle/fno_port.py: rfftn + einsum + irfftn, /root/reference/Models/fno/spectral_layers.py:84-115,
183-208, 299-338) evaluated in float64 ON THE GPU (same torch ops, so it finishes in well under a second per case).
The reduced-batch cases of test_gpu_parity.py never fill the batch tiles of bypass_wgrad_tc (4 stacked batch entries
per MMA) or mode_mix_bwd_tile (batch tiling); these do.  Also: the 10-step roll-out TRAINING step of
Demo/Turbulence_2d+t/run_FNO&UNet.py:61-75 (loss + every parameter gradient, BPTT through ten model calls).
Tolerance 1e-5 relative L2 (north_star)."""
import pytest
import torch

from golden_util import rel_l2

pytestmark = pytest.mark.gpu

TOL = 1e-5

BENCH_SHAPES = [
    # (id, B, C, spatial, modes)
    ("C1_burgers_B20", 20, 64, (1026,), (16,)),
    ("C2_darcy_B20", 20, 32, (94, 94), (12, 12)),
    ("C3_ns_B20", 20, 64, (72, 72), (12, 12)),
    ("C4b_pipe_B20", 20, 128, (137, 137), (12, 12)),
    ("C5_turb3d_B10", 10, 32, (70, 70, 46), (8, 8, 8)),
]


def _dev():
    return torch.device("cuda", 0)


@pytest.mark.parametrize("path", ["simt", "auto"])
@pytest.mark.parametrize("case", BENCH_SHAPES, ids=[c[0] for c in BENCH_SHAPES])
def test_full_batch_layer_gradients_at_bench_shapes(case, path, monkeypatch):
    from conftest import set_deno_env
    from deno4pytorch_b200.functional import spectral_conv
    from oracle import fno_port
    set_deno_env(monkeypatch, "DENO_SPECTRAL_PATH", path)
    _, B, C, spatial, modes = case
    nd = len(spatial)
    dev = _dev()
    gen = torch.Generator().manual_seed(17)
    x = torch.randn((B, C) + spatial, generator=gen)
    ws = [torch.view_as_complex(torch.randn((C, C) + modes + (2,), generator=gen) / C) for _ in range(2 ** (nd - 1))]
    lw = torch.randn((C, C) + (1,) * nd, generator=gen) / C ** 0.5
    lb = torch.randn(C, generator=gen)
    cot = torch.randn((B, C) + spatial, generator=gen)

    xg = x.to(dev).requires_grad_(True)
    wg = [w.to(dev).requires_grad_(True) for w in ws]
    lwg, lbg = lw.to(dev).requires_grad_(True), lb.to(dev).requires_grad_(True)
    y = spectral_conv(xg, wg, lwg, lbg, modes, "gelu")
    y.backward(cot.to(dev))
    torch.cuda.synchronize()

    xr = x.to(dev).double().requires_grad_(True)
    wr = [w.to(dev).to(torch.complex128).requires_grad_(True) for w in ws]
    lwr, lbr = lw.to(dev).double().requires_grad_(True), lb.to(dev).double().requires_grad_(True)
    ref = fno_port.spectral_conv(xr, wr, lwr, lbr, modes, activation="gelu")
    ref.backward(cot.to(dev).double())
    torch.cuda.synchronize()

    assert rel_l2(y.detach().cpu(), ref.detach().cpu()) < TOL
    assert rel_l2(xg.grad.cpu(), xr.grad.cpu()) < TOL
    for j, (a, b) in enumerate(zip(wg, wr)):
        assert rel_l2(a.grad.cpu(), b.grad.cpu()) < TOL, f"weights{j + 1}"
    assert rel_l2(lwg.grad.cpu(), lwr.grad.cpu()) < TOL
    assert rel_l2(lbg.grad.cpu(), lbr.grad.cpu()) < TOL


def _rollout_loss(forward, x, grid, target, T):
    """Demo/Turbulence_2d+t/run_FNO&UNet.py:61-75: T model calls, each prediction appended to the input window and the
    oldest frame dropped; ONE loss over the concatenated predictions; one backward through all calls."""
    xx = x
    preds = []
    for _ in range(T):
        im = forward(xx, grid)
        preds.append(im)
        xx = torch.cat((xx[..., 1:], im), dim=-1)
    pred = torch.cat(preds, dim=-1)
    return torch.nn.functional.mse_loss(pred, target)


@pytest.mark.parametrize("path", ["simt", "auto"])
def test_rollout_training_gradients_match_port(path, monkeypatch):
    """The C3 training step as BASELINE.json states it: FNO2d(in 10, width 64, modes 12, pad 8) on 64x64, ten
    autoregressive calls, BPTT.  Loss and ALL parameter gradients against the float64 port on the GPU.  Every
    parameter is used ten times per backward, so gradients go through autograd's accumulation (no direct writes)."""
    from conftest import set_deno_env
    import deno4pytorch_b200 as d4
    from oracle import fno_port
    set_deno_env(monkeypatch, "DENO_SPECTRAL_PATH", path)
    dev = _dev()
    torch.manual_seed(12)
    cfg = dict(in_dim=10, out_dim=1, modes=(12, 12), width=64, depth=4, hidden=128, steps=1, padding=8)
    T, B = 10, 4
    model = d4.FNO2d(**cfg)
    with torch.no_grad():                     # reference init makes the spectral branch ~1e-4 of the bypass: lift it
        for k, p in model.named_parameters():   # (x256; x1024 makes the untrained ten-step roll-out grow 10x per step,
            if ".weights" in k:                 # an ill-conditioned comparison in any precision)
                p.mul_(cfg["width"] ** 2 / 16)
    params = {k: v.detach().clone() for k, v in model.state_dict().items()}
    gen = torch.Generator().manual_seed(13)
    x = torch.randn(B, 64, 64, 10, generator=gen)
    target = torch.randn(B, 64, 64, T, generator=gen)
    grid = fno_port.make_grid(B, (64, 64))
    model = model.to(dev)
    loss = _rollout_loss(model, x.to(dev), grid.to(dev), target.to(dev), T)
    loss.backward()
    torch.cuda.synchronize()

    p64 = {k: (v.to(dev).to(torch.complex128) if v.is_complex() else v.to(dev).double()).requires_grad_(True)
           for k, v in params.items()}
    fwd = lambda a, g: fno_port.fno_forward(p64, a, g, ndim=2, modes=cfg["modes"], depth=cfg["depth"],
                                            padding=cfg["padding"])
    ref = _rollout_loss(fwd, x.to(dev).double(), grid.to(dev).double(), target.to(dev).double(), T)
    ref.backward()
    # The same port in fp32 on the same GPU (= the reference's own arithmetic, north_star's comparison target): its
    # distance from the float64 evaluation is the conditioning of each gradient.  fc0.bias is a sum of 163 840
    # cancelling contributions propagated through 40 layer applications (|grad| ~ 1e-5 against O(1) elsewhere); the
    # reference's fp32 path itself is 1e-5-ish from float64 there.  Bar per parameter: north_star's 1e-5, or twice the
    # reference's own fp32 error where that is larger.
    p32 = {k: v.to(dev).clone().requires_grad_(True) for k, v in params.items()}
    fwd32 = lambda a, g: fno_port.fno_forward(p32, a, g, ndim=2, modes=cfg["modes"], depth=cfg["depth"],
                                              padding=cfg["padding"])
    _rollout_loss(fwd32, x.to(dev), grid.to(dev), target.to(dev), T).backward()
    assert abs(loss.item() - ref.item()) / abs(ref.item()) < TOL
    for k, p in model.named_parameters():
        ref32_err = rel_l2(p32[k].grad.cpu(), p64[k].grad.cpu())
        assert rel_l2(p.grad.cpu(), p64[k].grad.cpu()) < max(TOL, 2.0 * ref32_err), (k, ref32_err)


def test_rollout_trainstep_tracks_port_adam():
    """TrainStep(forward_fn=roll-out): three Adam steps of the BPTT roll-out (eager and graph replay share
    _forward_backward) follow the float64 port's torch.optim.Adam trajectory."""
    import deno4pytorch_b200 as d4
    from deno4pytorch_b200.training import TrainStep
    from oracle import fno_port
    dev = _dev()
    torch.manual_seed(14)
    cfg = dict(in_dim=4, out_dim=1, modes=(6, 6), width=32, depth=2, hidden=64, steps=1, padding=4)
    T, B = 5, 3
    model = d4.FNO2d(**cfg)
    with torch.no_grad():
        for k, p in model.named_parameters():
            if ".weights" in k:
                p.mul_(cfg["width"] ** 2 / 4)
    params = {k: v.detach().clone() for k, v in model.state_dict().items()}
    gen = torch.Generator().manual_seed(15)
    xs = [torch.randn(B, 28, 28, 4, generator=gen) for _ in range(3)]
    ts = [torch.randn(B, 28, 28, T, generator=gen) for _ in range(3)]
    grid = fno_port.make_grid(B, (28, 28))
    model = model.to(dev)
    fwd_fn = lambda m, a, g, t: _rollout_loss(m, a, g, t, T)
    step = TrainStep(model, xs[0].to(dev), grid.to(dev), ts[0].to(dev), forward_fn=fwd_fn, use_graph=True, warmup=2)
    # the constructor's warm-up iterations must leave parameters and optimiser untouched
    for k, v in model.state_dict().items():
        assert torch.equal(v.cpu(), params[k]), k
    assert float(step.opt.step_t) == 0.0 and float(step.opt.exp_avg.abs().max()) == 0.0

    p64 = {k: (v.to(dev).to(torch.complex128) if v.is_complex() else v.to(dev).double()).requires_grad_(True)
           for k, v in params.items()}
    opt = torch.optim.Adam(list(p64.values()), lr=1e-3, betas=(0.7, 0.9), weight_decay=1e-4)
    fwd = lambda a, g: fno_port.fno_forward(p64, a, g, ndim=2, modes=cfg["modes"], depth=cfg["depth"],
                                            padding=cfg["padding"])
    for x, t in zip(xs, ts):
        ref = _rollout_loss(fwd, x.to(dev).double(), grid.to(dev).double(), t.to(dev).double(), T)
        opt.zero_grad()
        ref.backward()
        opt.step()
        loss = float(step(x.to(dev), grid.to(dev), t.to(dev)))
        assert abs(loss - ref.item()) / abs(ref.item()) < TOL
    for k, p in model.named_parameters():
        assert rel_l2(p.detach().cpu(), p64[k].detach().cpu()) < TOL, k


def test_no_grad_forward_saves_nothing_and_matches():
    """Under torch.no_grad() (validation, roll-out inference) the layer neither allocates nor writes act'(z) / the saved
    spectrum and returns the same numbers."""
    from deno4pytorch_b200.functional import spectral_conv
    dev = _dev()
    gen = torch.Generator().manual_seed(2)
    x = torch.randn(2, 8, 30, 28, generator=gen).to(dev)
    ws = [torch.view_as_complex(torch.randn(8, 8, 5, 6, 2, generator=gen) / 8).to(dev).requires_grad_(True) for _ in range(2)]
    lw = (torch.randn(8, 8, 1, 1, generator=gen) / 8 ** 0.5).to(dev).requires_grad_(True)
    y1 = spectral_conv(x, ws, lw, None, (5, 6), "gelu")
    assert y1.grad_fn is not None and len(y1.grad_fn.saved_tensors) > 0
    torch.cuda.synchronize()
    base = torch.cuda.memory_allocated(dev)
    with torch.no_grad():
        y2 = spectral_conv(x, ws, lw, None, (5, 6), "gelu")
    torch.cuda.synchronize()
    assert y2.grad_fn is None
    assert torch.cuda.memory_allocated(dev) - base <= y2.numel() * 4 + 4096      # only the output stays alive
    assert torch.equal(y1.detach(), y2)


def test_second_backward_on_a_trainstep_model_accumulates_correctly():
    """direct gradient writes overwrite.  A parameter used in two live graphs (gradient accumulation, a loss
    that calls the model twice) must fall back to autograd's accumulation; after TrainStep.close() the sinks are gone."""
    import deno4pytorch_b200 as d4
    from deno4pytorch_b200.training import TrainStep
    dev = _dev()
    torch.manual_seed(3)
    model = d4.FNO2d(in_dim=1, out_dim=1, modes=(6, 6), width=32, depth=2, padding=3).to(dev)
    x = torch.randn(4, 29, 31, 1, device=dev)
    grid = torch.rand(4, 29, 31, 2, device=dev)
    tgt = torch.randn(4, 29, 31, 1, device=dev)
    step = TrainStep(model, x, grid, tgt, use_graph=False, warmup=1)
    assert all(getattr(p, "_deno_grad_sink", None) is not None for p in model.parameters())

    def grads_of(fn):
        step.grads.zero()
        step.grads.reset_pending()
        fn().backward()
        torch.cuda.synchronize()
        return step.grads.flat.clone()

    mse = torch.nn.functional.mse_loss
    once = grads_of(lambda: mse(model(x, grid), tgt))
    twice = grads_of(lambda: mse(model(x, grid), tgt) + mse(model(x, grid), tgt))       # two live uses of every parameter
    assert rel_l2(twice.cpu(), (2.0 * once).cpu()) < 1e-6
    step.close()
    assert all(getattr(p, "_deno_grad_sink", None) is None for p in model.parameters())
    step.grads.zero()
    mse(model(x, grid), tgt).backward()
    mse(model(x, grid), tgt).backward()                                                 # accumulation over two backwards
    torch.cuda.synchronize()
    assert rel_l2(step.grads.flat.cpu(), (2.0 * once).cpu()) < 1e-6
