#!/bin/bash
# Session-3 visit P: mode_mix_lane A/B (DENO_MIX_LANE = 0 / 1 / 2) + parity of the lane kernel on the GPU.
tag=${1:-r3p}
mkdir -p gpurun_out
DENO_MIX_LANE=2 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullbatch.py -m gpu -q --maxfail=5 -p no:cacheprovider > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
grep -E "^FAILED|passed|failed|exit|Error" gpurun_out/${tag}_pytest.log | tail -8
timeout 900 python tools/ab_env.py "DENO_MIX_LANE=0" "DENO_MIX_LANE=2" -- C3 C2 C4b C5 C1 > gpurun_out/${tag}_ab.log 2>&1
cat gpurun_out/${tag}_ab.log
