#!/bin/bash
# Session-3 visit I: row_synthesis_tc - parity (whole GPU suite) + A/B against DENO_ROWSYNTH_TC=0.
tag=${1:-r3i}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
grep -E "^FAILED|passed|failed|exit|Error" gpurun_out/${tag}_pytest.log | tail -15
timeout 900 python tools/ab_env.py "DENO_ROWSYNTH_TC=0" "DENO_ROWSYNTH_TC=1" -- C2 C3 C5 C4b > gpurun_out/${tag}_ab.log 2>&1
cat gpurun_out/${tag}_ab.log
