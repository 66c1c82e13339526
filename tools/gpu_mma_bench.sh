#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./tools/_build/tc_mma_bench > gpurun_out/tc_mma_bench.log 2>&1
echo "exit $?" >> gpurun_out/tc_mma_bench.log
cat gpurun_out/tc_mma_bench.log
