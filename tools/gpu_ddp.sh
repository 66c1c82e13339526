#!/bin/bash
# two-GPU DP check + scaling bench.  Usage: gpurun --gpus 2 -- bash tools/gpu_ddp.sh tag
tag=${1:-ddp}
mkdir -p gpurun_out
timeout ${TMO:-120} python -m pytest tests/test_gpu_ddp.py -q -x -m gpu -p no:cacheprovider > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -n 15 gpurun_out/${tag}_pytest.log
bash tools/gpu_scale.sh 2 ${tag}
