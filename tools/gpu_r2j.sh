#!/bin/bash
# Round-2 visit J: tensor-core channel mix (parity at the config / full-batch shapes), TMA analysis after the extent fix.
tag=${1:-r2j}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullbatch.py -m gpu -q --maxfail=30 -p no:cacheprovider -k "config_shapes or full_batch or rollout" > gpurun_out/${tag}_pytest_mix.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest_mix.log
grep -E "^FAILED|passed|failed|exit" gpurun_out/${tag}_pytest_mix.log | tail -20
grep -E "^E  +assert" gpurun_out/${tag}_pytest_mix.log | cut -c1-200 | head -20
for w in ns2d_w64 pipe2d_w128; do
  for m in 1 0; do
  DENO_MIX_TC=$m timeout 600 python bench.py --workload $w --steps 20 --no-secondary --cpu-steps 1 > gpurun_out/${tag}_bench_${w}_m$m.log 2>&1
  echo "bench $w mix_tc=$m exit $?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_bench_${w}_m$m.log").read().strip().splitlines()[-1])
    print(" value %.1f ms %.3f layer_us %.1f frac %.3f vs_torch %.2f" % (d["value"], d["ms_per_step"], d["roofline"]["layer_fwd_bwd"]["us"], d["roofline"]["layer_fwd_bwd"]["frac"], d.get("vs_torch_gpu") or 0))
    print(" kernels", d["roofline"]["kernels_us"])
except Exception as e:
    print(" parse failed", e)
PY
  done
done
DENO_ANALYSIS_TMA=1 timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/${tag}_pytest_antma.log 2>&1
echo "pytest antma exit $?" >> gpurun_out/${tag}_pytest_antma.log
grep -E "^FAILED|passed|failed|exit" gpurun_out/${tag}_pytest_antma.log | tail -8
