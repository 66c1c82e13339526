#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3 4 5; do
  timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -4 > gpurun_out/flaky_$i.log
  grep -E "passed|failed" gpurun_out/flaky_$i.log
  grep -E "^FAILED" gpurun_out/flaky_$i.log
done
