#!/bin/bash
# ncu --set full capture of ONE kernel family from the layer probe.  Usage: gpurun -- bash tools/gpu_ncu_one.sh <regex> <tag> [skip]
rx=${1:-row_synthesis}; tag=${2:-one}; skip=${3:-4}
mkdir -p gpurun_out
python tools/layer_probe.py C2 3 > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"$rx" --launch-skip $skip --launch-count 1 \
    -o gpurun_out/${tag} -f python tools/layer_probe.py C2 3 > gpurun_out/${tag}_ncu.log 2>&1
echo "ncu exit $?" >> gpurun_out/${tag}_ncu.log
tail -n 3 gpurun_out/${tag}_ncu.log
