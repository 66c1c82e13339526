#!/bin/bash
mkdir -p gpurun_out
for ws in 0 1; do
  DENO_SYNTH_WS=$ws timeout 600 python bench.py --workload ns2d_w64 --steps 20 --warmup 3 --cpu-steps 1 > gpurun_out/ns_ws$ws.log 2>&1
  DENO_SYNTH_WS=$ws timeout 600 python bench.py --workload turb3d --steps 10 --warmup 3 --cpu-steps 1 > gpurun_out/t3_ws$ws.log 2>&1
done
echo done
