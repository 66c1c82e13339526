#!/bin/bash
tag=${1:-run}
mkdir -p gpurun_out
timeout 120 ./tools/_build/tc_mma_bench > gpurun_out/${tag}_mma_bench.log 2>&1
bash tools/gpu_check.sh $tag
