#!/bin/bash
# Session-3 visit U: gz product fused into the transform loop, 64-bit staged-plane accesses - parity + layer / step times.
tag=${1:-r3u}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullbatch.py -m gpu -q --maxfail=5 -p no:cacheprovider > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
grep -E "^FAILED|passed|failed|exit|Error" gpurun_out/${tag}_pytest.log | tail -8
timeout 900 python tools/ab_env.py "X=now" -- C2 C3 C5 C4b > gpurun_out/${tag}_ab.log 2>&1
cat gpurun_out/${tag}_ab.log
