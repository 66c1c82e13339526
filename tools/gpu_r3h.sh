#!/bin/bash
# Session-3 visit H: analysis MMA-first reorder A/B vs ring-off; ncu --set full of mode_mix_tc and row_synthesis at C3.
tag=${1:-r3h}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullbatch.py -m gpu -q -x -p no:cacheprovider > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
timeout 600 python tools/ab_env.py "DENO_AN_RING=0" "DENO_AN_RING=1" -- C2 C3 C5 > gpurun_out/${tag}_ab.log 2>&1
cat gpurun_out/${tag}_ab.log
for rx in mode_mix_tc row_synthesis; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$rx" --launch-skip 2 --launch-count 1 \
    -o gpurun_out/${tag}_C3_${rx} -f python tools/layer_probe.py C3 3 > gpurun_out/${tag}_ncu_${rx}.log 2>&1
  echo "ncu exit $?" >> gpurun_out/${tag}_ncu_${rx}.log
  tail -n 2 gpurun_out/${tag}_ncu_${rx}.log
done
