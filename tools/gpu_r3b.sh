#!/bin/bash
# Session-3 visit B: backward overlap (parameter-gradient kernels on a side stream) - parity + A/B.
tag=${1:-r3b}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "overlap or repeat or direct_gradient or trainstep" -p no:cacheprovider > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
timeout 900 python tools/ab_env.py "DENO_BWD_OVERLAP=0" "DENO_BWD_OVERLAP=1" -- C2 C3 C1 C5 C4b > gpurun_out/${tag}_ab.log 2>&1
cat gpurun_out/${tag}_ab.log
