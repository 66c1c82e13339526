#!/bin/bash
# Round-2 visit B: the TMA kernels.  Parity suite with each new kernel on its own, then both; smoke; bench A/B.
tag=${1:-r2b}
mkdir -p gpurun_out
run_suite() {  # name, env...
  local name=$1; shift
  env "$@" timeout 300 python -m pytest tests -m gpu -q --maxfail=8 -p no:cacheprovider > gpurun_out/${tag}_pytest_${name}.log 2>&1
  echo "pytest[$name] exit $?" | tee -a gpurun_out/${tag}_pytest_${name}.log
  grep -E "passed|failed" gpurun_out/${tag}_pytest_${name}.log | tail -2
  grep -E "^FAILED" gpurun_out/${tag}_pytest_${name}.log | head -12
}
run_suite wgrad DENO_ANALYSIS_TMA=0 DENO_WGRAD_TMA=1
run_suite analysis DENO_ANALYSIS_TMA=1 DENO_WGRAD_TMA=0
run_suite both DENO_ANALYSIS_TMA=1 DENO_WGRAD_TMA=1
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/${tag}_smoke.log
tail -2 gpurun_out/${tag}_smoke.log
for v in "0 0" "0 1" "1 0" "1 1"; do
  set -- $v
  DENO_ANALYSIS_TMA=$1 DENO_WGRAD_TMA=$2 timeout 300 python bench.py --steps 100 --no-secondary --cpu-steps 2 > gpurun_out/${tag}_bench_a$1_w$2.log 2>&1
  echo "bench a=$1 w=$2 exit $?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_bench_a$1_w$2.log").read().strip().splitlines()[-1])
    print(" value %.0f e2e %.0f ms %.3f layer_us %.1f frac %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["layer_fwd_bwd"]["us"], d["roofline"]["layer_fwd_bwd"]["frac"]))
    print(" kernels", d["roofline"]["kernels_us"])
except Exception as e:
    print(" parse failed", e)
PY
done
timeout 600 python bench.py > gpurun_out/${tag}_bench_full.log 2>&1
echo "bench full exit $?"
tail -c 2500 gpurun_out/${tag}_bench_full.log
