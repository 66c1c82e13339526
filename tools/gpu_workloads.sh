#!/bin/bash
# bench.py on the other BASELINE.json configs (secondary numbers for DESIGN.md).
mkdir -p gpurun_out
for w in burgers1d ns2d_w64 turb3d pipe2d_w128; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --cpu-steps 2 > gpurun_out/wl_$w.log 2>&1
  echo "$w exit $?"; tail -c 300 gpurun_out/wl_$w.log | head -c 300; echo
done
