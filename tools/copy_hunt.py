"""Where do activation-sized copies come from in one training step of a workload?  torch profiler, shapes + stacks.
python tools/copy_hunt.py pipe2d_w128"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

wl = sys.argv[1] if len(sys.argv) > 1 else "pipe2d_w128"
dev = torch.device("cuda", 0)
import deno4pytorch_b200 as d4
ndim, kw, spatial, batch = bench.WORKLOADS[wl]
torch.manual_seed(0)
model = {1: d4.FNO1d, 2: d4.FNO2d, 3: d4.FNO3d}[ndim](**kw).to(dev)
x, grid, target = bench.workload_batch(wl, seed=1)
x, grid, target = x.to(dev), grid.to(dev), target.to(dev)
for _ in range(2):
    model.zero_grad()
    torch.nn.functional.mse_loss(model(x, grid), target).backward()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True) as prof:
    model.zero_grad()
    torch.nn.functional.mse_loss(model(x, grid), target).backward()
    torch.cuda.synchronize()
for ev in prof.events():
    if ev.name in ("aten::copy_", "aten::contiguous", "aten::clone") and ev.device_time_total > 20:
        print(ev.name, "cuda_us=%.0f" % ev.device_time_total, ev.input_shapes, [str(f) for f in (ev.stack or [])[:6]])
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
