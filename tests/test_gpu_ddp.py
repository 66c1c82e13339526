"""Two-GPU checks of the data-parallel training step (need >= 2 GPUs: skipped on the one-GPU box the suite normally runs
on; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_ddp.py -m gpu`).

Semantics under test (/root/reference/Demo/Turbulence_3d+t/run_train_DDP.py:64-101,272: DistributedDataParallel averages
gradients over ranks): two ranks, each on its half of a batch, must end a step holding EXACTLY the gradient one rank
computes on the whole batch (MSE is a mean, so the mean of the two half-batch gradients is the full-batch gradient), to
fp32 reduction order: <= 1e-5 relative L2 per parameter, no optimiser in between (lr = 0, weight decay 0 keeps the
parameters fixed while the whole captured step - forward, backward, collective, optimiser kernel - still runs).
Every collective flavour TrainStep offers is checked: the averaged NCCL all-reduce captured inside the step graph (the
default SCALE runs), the per-stage overlapped all-reduce, NCCL between two graphs, and the fused
all-reduce + Adam kernel over NVLink multicast (``DENO_DP_FUSED=1``).  A second test checks that with a real learning
rate the two-rank trajectory equals the single-rank full-batch trajectory."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu

MODES = {
    "ingraph": dict(DENO_DP_OVERLAP="0", DENO_DP_INGRAPH="1", DENO_DP_FUSED="0", DENO_DP_FUSED_WRITE_GRADS="0"),
    "overlap": dict(DENO_DP_OVERLAP="1", DENO_DP_INGRAPH="1", DENO_DP_FUSED="0", DENO_DP_FUSED_WRITE_GRADS="0"),
    "two_graphs": dict(DENO_DP_OVERLAP="0", DENO_DP_INGRAPH="0", DENO_DP_FUSED="0", DENO_DP_FUSED_WRITE_GRADS="0"),
    # the fused kernel consumes the reduced gradient in registers; the test asks it to store the average back as well
    "fused": dict(DENO_DP_OVERLAP="0", DENO_DP_INGRAPH="1", DENO_DP_FUSED="1", DENO_DP_FUSED_WRITE_GRADS="1"),
}


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make(dev):
    import deno4pytorch_b200 as d4
    torch.manual_seed(5)
    model = d4.FNO2d(in_dim=1, out_dim=1, modes=(6, 6), width=32, depth=3, padding=3)
    with torch.no_grad():                       # spectral branch O(1) of the bypass, so its gradients matter in the norm
        for k, p in model.named_parameters():
            if ".weights" in k:
                p.mul_(32 ** 2 / 4)
    model = model.to(dev)
    g = torch.Generator().manual_seed(6)
    x = torch.randn(4, 29, 31, 1, generator=g)
    grid = torch.rand(4, 29, 31, 2, generator=g)
    tgt = torch.randn(4, 29, 31, 1, generator=g)
    return model, x, grid, tgt


def _worker(rank, world, port, out_path, env, lr, nsteps):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), **env)
    import faulthandler
    import sys
    faulthandler.dump_traceback_later(150, exit=True, file=sys.stderr)      # a hung rank ends the test, not the box
    from deno4pytorch_b200.training import TrainStep
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    model, x, grid, tgt = _make(dev)
    per = x.shape[0] // world
    sl = slice(rank * per, (rank + 1) * per)
    step = TrainStep(model, x[sl].to(dev), grid[sl].to(dev), tgt[sl].to(dev), lr=lr, weight_decay=0.0 if lr == 0 else 1e-4,
                     warmup=2, overlap_allreduce=env["DENO_DP_OVERLAP"] == "1")
    assert step._overlap == (env["DENO_DP_OVERLAP"] == "1")
    assert step._fused_dp == (env["DENO_DP_FUSED"] == "1"), "fused all-reduce + Adam path was not taken"
    losses = [float(step()) for _ in range(nsteps)]
    torch.cuda.synchronize()
    if rank == 0:
        torch.save({"losses": losses, "grads": step.grads.flat.detach().cpu(),
                    "end": {k: v.detach().cpu() for k, v in model.state_dict().items()}}, out_path)
    step.close()
    dist.barrier()
    dist.destroy_process_group()


def _spawn(tmp_path, mode, lr, nsteps):
    import torch.multiprocessing as mp
    out = str(tmp_path / f"ddp_{mode}.pt")
    mp.spawn(_worker, args=(2, _free_port(), out, MODES[mode], lr, nsteps), nprocs=2, join=True)
    return torch.load(out)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("mode", list(MODES))
def test_two_rank_averaged_gradients_equal_single_rank_full_batch(tmp_path, mode):
    from deno4pytorch_b200.training import TrainStep
    from golden_util import rel_l2
    got = _spawn(tmp_path, mode, 0.0, 2)
    dev = torch.device("cuda", 0)
    model, x, grid, tgt = _make(dev)
    step = TrainStep(model, x.to(dev), grid.to(dev), tgt.to(dev), lr=0.0, weight_decay=0.0, warmup=1)
    step()
    torch.cuda.synchronize()
    ref = step.grads.flat.detach().cpu()
    assert got["grads"].shape == ref.shape
    assert rel_l2(got["grads"], ref) < 1e-5
    for p in step.grads.params:                                   # and per parameter
        off, n = step.grads.spans[id(p)]
        assert rel_l2(got["grads"][off:off + n], ref[off:off + n]) < 1e-5
    for k, v in model.state_dict().items():                       # lr = 0: nothing moved on either side
        assert torch.equal(got["end"][k], v.detach().cpu()), k


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("mode", ["ingraph", "fused"])
def test_two_rank_trajectory_equals_single_rank_full_batch(tmp_path, mode):
    """Same init (TrainStep's constructor leaves parameters and optimiser state untouched), same data, three Adam steps:
    parameters agree to fp32 reduction order."""
    from deno4pytorch_b200.training import TrainStep
    from golden_util import rel_l2
    got = _spawn(tmp_path, mode, 1e-3, 3)
    dev = torch.device("cuda", 0)
    model, x, grid, tgt = _make(dev)
    step = TrainStep(model, x.to(dev), grid.to(dev), tgt.to(dev), lr=1e-3, warmup=1)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    for k, v in model.state_dict().items():
        a, b = got["end"][k], v.detach().cpu()
        a = torch.view_as_real(a) if a.is_complex() else a
        b = torch.view_as_real(b) if b.is_complex() else b
        assert rel_l2(a, b) < 1e-5, k
