#!/bin/bash
tag=${1:-run}
bash tools/gpu_check.sh $tag
cd tools && timeout 120 python phase_probe.py C2 > ../gpurun_out/${tag}_phase.log 2>&1; cd ..
tail -3 gpurun_out/${tag}_phase.log
