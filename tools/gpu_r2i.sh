#!/bin/bash
# Round-2 visit I: TMA analysis kernel with per-class shifted twiddle images.
tag=${1:-r2i}
mkdir -p gpurun_out
timeout 600 python tools/an_tma_debug.py quick > gpurun_out/${tag}_an_tma_debug.log 2>&1
echo "debug exit $?" >> gpurun_out/${tag}_an_tma_debug.log
cat gpurun_out/${tag}_an_tma_debug.log
DENO_ANALYSIS_TMA=1 timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/${tag}_pytest_antma.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest_antma.log
grep -E "^FAILED|passed|failed|exit" gpurun_out/${tag}_pytest_antma.log | tail -14
for a in 0 1; do
  DENO_ANALYSIS_TMA=$a timeout 300 python bench.py --steps 100 --no-secondary --cpu-steps 1 > gpurun_out/${tag}_bench_a$a.log 2>&1
  echo "bench a=$a exit $?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_bench_a$a.log").read().strip().splitlines()[-1])
    print(" value %.0f e2e %.0f ms %.3f layer_us %.1f frac %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["layer_fwd_bwd"]["us"], d["roofline"]["layer_fwd_bwd"]["frac"]))
    print(" kernels", d["roofline"]["kernels_us"])
except Exception as e:
    print(" parse failed", e)
PY
done
