#!/bin/bash
# Session-3 visit Q: lane-mix unroll 16 A/B at C3 / C4b, attribution of every workload's layer, full suite, full default bench.
tag=${1:-r3q}
mkdir -p gpurun_out
timeout 600 python tools/ab_env.py "DENO_MIX_LANE=0" "DENO_MIX_LANE=1" -- C3 C4b > gpurun_out/${tag}_ab.log 2>&1
cat gpurun_out/${tag}_ab.log
timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
grep -E "^FAILED|passed|failed|exit|Error" gpurun_out/${tag}_pytest.log | tail -8
( time timeout 900 python bench.py --gpus 1 --steps 200 --warmup 5 ) > gpurun_out/${tag}_bench.log 2>&1
echo "bench exit $?" >> gpurun_out/${tag}_bench.log
tail -c 400 gpurun_out/${tag}_bench.log
