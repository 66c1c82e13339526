#!/bin/bash
mkdir -p gpurun_out
cd tools && timeout 120 python phase_probe.py C2 > ../gpurun_out/phase_probe.log 2>&1; cd ..
cat gpurun_out/phase_probe.log | tail -12
