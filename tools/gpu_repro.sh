#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/repro_case.py > gpurun_out/repro.log 2>&1; cat gpurun_out/repro.log | tail -8
