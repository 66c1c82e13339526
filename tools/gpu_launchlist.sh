#!/bin/bash
# Launch list of the bench command (every kernel with its device time): compare SHARES, not absolutes.
tag=${1:-ll}
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --cpu-steps 1 > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip ${SKIP:-540} --launch-count ${COUNT:-200} --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --cpu-steps 1 > gpurun_out/${tag}_ncu.log 2>&1
echo "exit $?" >> gpurun_out/${tag}_ncu.log
tail -c 600 gpurun_out/${tag}_plain.log; tail -n 3 gpurun_out/${tag}_ncu.log; wc -l gpurun_out/${tag}_launches.csv
