#!/bin/bash
# Final N=1 evidence of the round: GPU suite, smoke, full default bench, ncu --set full of one Darcy layer, launch list of one step.
tag=${1:-r4k}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
grep -E "^FAILED|passed|failed|exit" gpurun_out/${tag}_pytest.log | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/${tag}_smoke.log
tail -2 gpurun_out/${tag}_smoke.log
( time timeout 900 python bench.py --gpus 1 --steps 200 --warmup 5 ) > gpurun_out/${tag}_bench.log 2>&1
echo "bench exit $?" >> gpurun_out/${tag}_bench.log
tail -c 300 gpurun_out/${tag}_bench.log
( time timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 ) > gpurun_out/${tag}_bench_ref.log 2>&1
tail -c 400 gpurun_out/${tag}_bench_ref.log
SKIP=11 COUNT=11 bash tools/gpu_ncu.sh C2 ${tag} > gpurun_out/${tag}_ncu_driver.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 560 --launch-count 180 --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --cpu-steps 1 --no-secondary > gpurun_out/${tag}_ll.log 2>&1
echo "launchlist exit $?"; wc -l gpurun_out/${tag}_launches.csv
