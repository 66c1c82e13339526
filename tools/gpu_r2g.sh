#!/bin/bash
# Round-2 visit G: TMA box legality probe; wide-layer (channel tiles / K passes) synthesis parity; C4b bench.
tag=${1:-r2g}
mkdir -p gpurun_out
P=tools/_build/tma_box_probe
{
probe() { echo -n "inner=$1 rows=$2 row_bytes=$3 box_rows=$4 base_off=$5 c0=$6 c1=$7 : "; timeout 30 $P "$@" 2>&1 | tail -1; }
probe 16 4096 64 16 0 0 0
probe 94 4096 752 47 0 0 0
probe 94 4096 752 48 0 0 0
probe 94 4096 752 32 0 0 0
probe 96 4096 752 47 368 2 0
probe 96 4096 752 48 368 2 0
probe 94 4096 752 47 0 64 0
probe 94 4096 752 47 0 64 47
probe 96 4096 768 47 0 0 0
probe 96 4096 752 47 0 0 0
probe 92 4096 752 47 0 0 0
probe 94 4096 752 8 0 0 0
probe 94 4096 752 40 0 0 0
probe 94 4096 376 47 0 0 0
probe 8836 640 35344 32 0 8832 0
probe 72 4096 288 72 0 0 0
probe 72 4096 288 72 0 64 0
} > gpurun_out/${tag}_box_probe.log 2>&1
cat gpurun_out/${tag}_box_probe.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullbatch.py -m gpu -q --maxfail=30 -p no:cacheprovider -k "config_shapes or full_batch" > gpurun_out/${tag}_pytest_wide.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest_wide.log
grep -E "^FAILED|passed|failed|exit" gpurun_out/${tag}_pytest_wide.log | tail -20
timeout 600 python bench.py --workload pipe2d_w128 --steps 20 --no-secondary --cpu-steps 0 > gpurun_out/${tag}_bench_c4b.log 2>&1
echo "bench exit $?"
tail -c 2500 gpurun_out/${tag}_bench_c4b.log
