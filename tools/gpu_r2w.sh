#!/bin/bash
# Round-2 visit W: A/B vs the session-start build after the compile-time special cases; suite.
tag=${1:-r2w}
mkdir -p gpurun_out
python tools/ab_layer.py C5 C2 C3 C1 > gpurun_out/${tag}_ab.log 2>&1
cat gpurun_out/${tag}_ab.log
timeout 900 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
grep -E "^FAILED|passed|failed|exit" gpurun_out/${tag}_pytest.log | tail -12
