#!/bin/bash
# Round-2 session-3 visit A: suite, smoke, full default bench (driver's command) on the restored build.
tag=${1:-r3a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${tag}_gpu.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/${tag}_smoke.log
( time timeout 900 python bench.py --gpus 1 --steps 200 --warmup 5 ) > gpurun_out/${tag}_bench.log 2>&1
echo "bench exit $?" >> gpurun_out/${tag}_bench.log
grep -E "^FAILED|passed|failed|exit" gpurun_out/${tag}_pytest.log | tail -12; tail -3 gpurun_out/${tag}_smoke.log; tail -c 1500 gpurun_out/${tag}_bench.log
