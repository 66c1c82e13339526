#!/bin/bash
# ncu --set full of one Darcy layer + launch list of one training step on the final build.
tag=${1:-r4p}
mkdir -p gpurun_out
SKIP=11 COUNT=11 bash tools/gpu_ncu.sh C2 ${tag} > gpurun_out/${tag}_ncu_driver.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 560 --launch-count 180 --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --cpu-steps 1 --no-secondary > gpurun_out/${tag}_ll.log 2>&1
echo "launchlist exit $?"; wc -l gpurun_out/${tag}_launches.csv
