#!/bin/bash
# Session-3 visit: 2-GPU DP tests + N=2 bench (fused multicast exchange and NCCL beside it).
tag=${1:-r4i}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ddp.py -q -m gpu -p no:cacheprovider > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -n 6 gpurun_out/${tag}_pytest.log
TMO=300 bash tools/gpu_scale2.sh 2 ${tag}
