#!/bin/bash
# One GPU-box visit: parity tests, smoke, short bench.  Usage (from the repo root, under gpurun):
#   gpurun --timeout 1500 -- bash tools/gpu_check.sh [tag]
# Everything lands in gpurun_out/<tag>_*.log
tag=${1:-run}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${tag}_gpu.log 2>&1
nproc >> gpurun_out/${tag}_gpu.log
timeout 900 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/${tag}_smoke.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/${tag}_bench.log 2>&1
echo "bench exit $?" >> gpurun_out/${tag}_bench.log
tail -5 gpurun_out/${tag}_pytest.log; tail -3 gpurun_out/${tag}_smoke.log; tail -3 gpurun_out/${tag}_bench.log
