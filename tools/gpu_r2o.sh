#!/bin/bash
# Round-2 visit O: ncu --set full of one layer fwd+bwd at C2 and C3 (third iteration), raw pages exported as CSV.
tag=${1:-r2o}
mkdir -p gpurun_out
for cfg in C2 C3; do
  python tools/layer_probe.py $cfg 3 > gpurun_out/${tag}_${cfg}_plain.log 2>&1 || { echo "probe $cfg failed"; continue; }
  # launches per iteration are at most 11; skip the first two iterations
  n=$(python - <<PY
print({"C2": 10, "C3": 11}["$cfg"])
PY
)
  ncu --set full --clock-control none --import-source on -k regex:'plane_|mode_|bypass_|outer_|row_|line_' \
      --launch-skip $((2 * n)) --launch-count $n -o gpurun_out/${tag}_${cfg} -f python tools/layer_probe.py $cfg 3 > gpurun_out/${tag}_${cfg}_ncu.log 2>&1
  echo "ncu $cfg exit $?"
  ncu -i gpurun_out/${tag}_${cfg}.ncu-rep --page raw --csv > gpurun_out/${tag}_${cfg}_raw.csv 2>/dev/null
  wc -l gpurun_out/${tag}_${cfg}_raw.csv
done
# launch list of one C4b step (which launches make up the 13.5 ms?)
python bench.py --workload pipe2d_w128 --steps 2 --warmup 3 --cpu-steps 1 --no-secondary > gpurun_out/${tag}_c4b_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 300 --launch-count 400 --csv \
    --log-file gpurun_out/${tag}_c4b_launches.csv python bench.py --workload pipe2d_w128 --steps 2 --warmup 3 --cpu-steps 1 --no-secondary > gpurun_out/${tag}_c4b_ncu.log 2>&1
echo "launch list exit $?"
python tools/launchlist_summary.py gpurun_out/${tag}_c4b_launches.csv | head -40
