#!/bin/bash
# Session-3 visit Y: ring sized by the rows that exist (the Darcy plane now gets it) - parity + A/B at C2 / C3 / C5.
tag=${1:-r3y}
mkdir -p gpurun_out
DENO_AN_DEBUG=1 python tools/layer_probe.py C2 1 2>&1 | grep "\[deno\]" | sort | uniq -c > gpurun_out/${tag}_cfg.log
cat gpurun_out/${tag}_cfg.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullbatch.py -m gpu -q --maxfail=5 -p no:cacheprovider > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
grep -E "^FAILED|passed|failed|exit|Error" gpurun_out/${tag}_pytest.log | tail -8
timeout 900 python tools/ab_env.py "DENO_AN_RING=0" "DENO_AN_RING=1" "DENO_AN_RING=2" -- C2 C3 C5 > gpurun_out/${tag}_ab.log 2>&1
cat gpurun_out/${tag}_ab.log
