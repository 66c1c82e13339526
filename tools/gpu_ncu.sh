#!/bin/bash
# ncu capture of one layer fwd+bwd (third iteration) at a config shape.  Usage under gpurun:
#   gpurun --timeout 900 -- bash tools/gpu_ncu.sh C2 tag
cfg=${1:-C2}; tag=${2:-ncu}
mkdir -p gpurun_out
python tools/layer_probe.py $cfg 3 > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'plane_|mode_|bypass_|outer_|row_' \
    --launch-skip ${SKIP:-16} --launch-count ${COUNT:-8} -o gpurun_out/${tag}_${cfg} -f python tools/layer_probe.py $cfg 3 > gpurun_out/${tag}_ncu.log 2>&1
echo "ncu exit $?" >> gpurun_out/${tag}_ncu.log
tail -3 gpurun_out/${tag}_plain.log gpurun_out/${tag}_ncu.log
