#!/bin/bash
# Session-3 visit C: analysis staging ring (two slots / staged act'(z)) - parity + A/B against DENO_AN_RING=0.
tag=${1:-r3c}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
timeout 900 python tools/ab_env.py "DENO_AN_RING=0" "DENO_AN_RING=1" "DENO_AN_RING=2" -- C2 C3 C5 C4b > gpurun_out/${tag}_ab.log 2>&1
cat gpurun_out/${tag}_ab.log
