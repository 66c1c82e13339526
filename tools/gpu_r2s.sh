#!/bin/bash
# Round-2 visit S: the committed profile evidence.  (1) launch list of the bench command (one C2 training step),
# (2) ncu --set full of one layer fwd+bwd at C2 and C3, exported as raw CSV (the .ncu-rep files stay on the box: > 64 MiB).
tag=${1:-r2s}
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --cpu-steps 1 --no-secondary > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip ${SKIP:-400} --launch-count ${COUNT:-260} --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --cpu-steps 1 --no-secondary > gpurun_out/${tag}_ncu_ll.log 2>&1
echo "launch list exit $?"
python tools/launchlist_summary.py gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_launch_list_one_step.txt
head -30 gpurun_out/${tag}_launch_list_one_step.txt
for cfg in C2 C3; do
  python tools/layer_probe.py $cfg 3 > gpurun_out/${tag}_${cfg}_plain.log 2>&1 || { echo "probe $cfg failed"; continue; }
  n=$(python -c "print({'C2': 10, 'C3': 11}['$cfg'])")
  ncu --set full --clock-control none -k regex:'plane_|mode_|bypass_|outer_|row_|line_' \
      --launch-skip $((2 * n)) --launch-count $n -o /tmp/${tag}_${cfg} -f python tools/layer_probe.py $cfg 3 > gpurun_out/${tag}_${cfg}_ncu.log 2>&1
  echo "ncu $cfg exit $?"
  ncu -i /tmp/${tag}_${cfg}.ncu-rep --page raw --csv > gpurun_out/${tag}_${cfg}_raw.csv 2>/dev/null
  wc -lc gpurun_out/${tag}_${cfg}_raw.csv
done
du -sh gpurun_out
