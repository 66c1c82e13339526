#!/bin/bash
# Round-2 visit A: full GPU suite, poisoned-buffer loops of the parity files (flake hunt), smoke, bench.
tag=${1:-r2a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${tag}_gpu.log 2>&1
nproc >> gpurun_out/${tag}_gpu.log
timeout 900 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -15 gpurun_out/${tag}_pytest.log
# poisoned loops: every output / scratch buffer NaN-filled before each library call; 4 processes x LOOPS suite passes
LOOPS=${LOOPS:-13}
for w in 1 2 3 4; do
  (
    for i in $(seq 1 $LOOPS); do
      DENO_POISON=1 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ends.py tests/test_gpu_regressor.py -m gpu -q -p no:cacheprovider 2>&1 | grep -E "passed|failed|FAILED|Error" | tr '\n' ' '
      echo
    done > gpurun_out/${tag}_poison_w$w.log 2>&1
  ) &
done
wait
cat gpurun_out/${tag}_poison_w*.log | sort | uniq -c | sort -rn | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/${tag}_smoke.log
tail -3 gpurun_out/${tag}_smoke.log
timeout 900 python bench.py > gpurun_out/${tag}_bench.log 2>&1
echo "bench exit $?" >> gpurun_out/${tag}_bench.log
tail -c 3000 gpurun_out/${tag}_bench.log
