#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./tools/_build/tc_layout_scan > gpurun_out/tc_layout_scan.log 2>&1
echo "exit $?" >> gpurun_out/tc_layout_scan.log
tail -5 gpurun_out/tc_layout_scan.log
