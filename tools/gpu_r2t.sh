#!/bin/bash
# Round-2 visit T: suite after the tall-plane (row slab) tensor-core analysis; C4b / C2 / C5 bench lines.
tag=${1:-r2t}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
grep -E "^FAILED|passed|failed|exit" gpurun_out/${tag}_pytest.log | tail -12
grep -E "^E  +(assert|Assertion|Runtime)" gpurun_out/${tag}_pytest.log | cut -c1-220 | head
for w in pipe2d_w128 darcy2d turb3d; do
  timeout 600 python bench.py --workload $w --steps 50 --no-secondary --cpu-steps 1 > gpurun_out/${tag}_bench_$w.log 2>&1
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
