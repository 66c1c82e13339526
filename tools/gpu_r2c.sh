#!/bin/bash
# Round-2 visit C: descriptor / TMA probe + structured wgrad diagnostic.
tag=${1:-r2c}
mkdir -p gpurun_out
timeout 60 tools/_build/tma_probe > gpurun_out/${tag}_tma_probe.log 2>&1
echo "probe exit $?" >> gpurun_out/${tag}_tma_probe.log
cat gpurun_out/${tag}_tma_probe.log
timeout 120 python tools/wgrad_diag.py > gpurun_out/${tag}_wgrad_diag.log 2>&1
echo "diag exit $?" >> gpurun_out/${tag}_wgrad_diag.log
cat gpurun_out/${tag}_wgrad_diag.log
