#!/bin/bash
# Session-3 visit O: per-kernel step profile (eager) + ncu --set full of the projection / lift kernels of one step.
tag=${1:-r3o}
mkdir -p gpurun_out
python tools/step_profile.py > gpurun_out/${tag}_step.log 2>&1
cat gpurun_out/${tag}_step.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fno_' --launch-skip 10 --launch-count 5 \
    -o gpurun_out/${tag}_ends -f python tools/step_profile.py > gpurun_out/${tag}_ncu.log 2>&1
echo "ncu exit $?" >> gpurun_out/${tag}_ncu.log
tail -n 2 gpurun_out/${tag}_ncu.log
