#!/bin/bash
tag=${1:-run}
mkdir -p gpurun_out
timeout 120 ./tools/_build/tc_probe > gpurun_out/${tag}_tcprobe.log 2>&1
echo "tc_probe exit $?" >> gpurun_out/${tag}_tcprobe.log
cat gpurun_out/${tag}_tcprobe.log
bash tools/gpu_check.sh $tag
