#!/bin/bash
# Round-2 visit N: wide-model ends test, C4b bench with the channels-first GEMM ends.
tag=${1:-r2n}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ends.py -m gpu -q --maxfail=30 -p no:cacheprovider > gpurun_out/${tag}_pytest_ends.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest_ends.log
grep -E "^FAILED|passed|failed|exit" gpurun_out/${tag}_pytest_ends.log | tail -12
grep -E "^E  +(assert|Assertion|Runtime)" gpurun_out/${tag}_pytest_ends.log | cut -c1-220 | head
for w in pipe2d_w128; do
  timeout 600 python bench.py --workload $w --steps 50 --no-secondary --cpu-steps 1 > gpurun_out/${tag}_bench_$w.log 2>&1
  echo "bench $w exit $?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_bench_$w.log").read().strip().splitlines()[-1])
    print(" value %.1f ms %.3f layer_us %.1f frac %.3f vs_torch %.2f launches/step %s" % (d["value"], d["ms_per_step"], d["roofline"]["layer_fwd_bwd"]["us"], d["roofline"]["layer_fwd_bwd"]["frac"], d.get("vs_torch_gpu") or 0, d.get("gpu_launches_per_step")))
except Exception as e:
    print(" parse failed", e)
PY
done
