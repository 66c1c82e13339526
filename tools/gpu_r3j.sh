#!/bin/bash
# Session-3 visit J: per-kernel attribution at C2/C5 and ncu --set full of row_synthesis_tc at C2.
tag=${1:-r3j}
mkdir -p gpurun_out
python - > gpurun_out/${tag}_attr.log 2>&1 <<'PY'
import sys, json, torch
sys.path.insert(0, '.')
import bench
for wl in ("darcy2d", "turb3d"):
    r = bench.layer_roofline(wl, torch.device("cuda", 0), 6539.9, reps=10, attribute=True)
    print(wl, round(r["layer_fwd_bwd"]["us"], 1), json.dumps(r["kernels_us"]))
PY
cat gpurun_out/${tag}_attr.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:row_synthesis_tc --launch-skip 2 --launch-count 1 \
    -o gpurun_out/${tag}_C2_rowsynth_tc -f python tools/layer_probe.py C2 3 > gpurun_out/${tag}_ncu.log 2>&1
echo "ncu exit $?" >> gpurun_out/${tag}_ncu.log
tail -n 2 gpurun_out/${tag}_ncu.log
