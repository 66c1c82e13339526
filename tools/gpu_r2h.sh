#!/bin/bash
# Round-2 visit H: which base-address alignment does a SWIZZLE_128B tensor map need?  + C4b / C3 bench lines.
tag=${1:-r2h}
mkdir -p gpurun_out
P=tools/_build/tma_box_probe
{
probe() { echo -n "inner=$1 rows=$2 row_bytes=$3 box_rows=$4 base_off=$5 c0=$6 c1=$7 : "; timeout 30 $P "$@" 2>&1 | tail -1; }
for off in 16 32 48 64 96 128 192 256 352 320 368 384; do probe 128 4096 752 47 $off 2 0; done
probe 128 4096 768 47 368 2 0
probe 128 4096 1024 47 16 0 0
} > gpurun_out/${tag}_box_probe.log 2>&1
cat gpurun_out/${tag}_box_probe.log
for w in pipe2d_w128 ns2d_rollout; do
  timeout 600 python bench.py --workload $w --steps 20 --no-secondary --cpu-steps 1 > gpurun_out/${tag}_bench_$w.log 2>&1
  echo "bench $w exit $?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_bench_$w.log").read().strip().splitlines()[-1])
    print(" value %.1f ms %.3f layer_us %.1f frac %.3f vs_torch %.2f" % (d["value"], d["ms_per_step"], d["roofline"]["layer_fwd_bwd"]["us"], d["roofline"]["layer_fwd_bwd"]["frac"], d.get("vs_torch_gpu") or 0))
    print(" kernels", d["roofline"]["kernels_us"])
except Exception as e:
    print(" parse failed", e)
PY
done
