#!/usr/bin/env python3
"""Condenses an `ncu -i X.ncu-rep --page raw --csv` export into one line per launch with the metrics DESIGN.md quotes:
duration, DRAM bytes read / written, SM and memory throughput (% of peak), warps active, tensor-pipe activity
(sm__pipe_tensor_cycles_active: the tensor-pipe utilisation north_star asks for on the GEMM paths), registers, grid.
    python tools/ncu_raw_summary.py gpurun_out/r2s_C2_raw.csv > profiles/round2_ncu_C2_layer_kernels.csv"""
import csv
import sys

COLS = [
    ("Kernel Name", "kernel"),
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_throughput_pct"),
    ("gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "mem_throughput_pct"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_throughput_pct"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_throughput_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor_pipe_active_pct"),
    ("sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "tmem_pipe_pct"),
    ("smsp__inst_executed.sum", "warp_instructions"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn_smem"),
]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
out = csv.writer(sys.stdout)
out.writerow([short + (" [" + units[idx[full]] + "]" if full in idx and units[idx[full]] else "") for full, short in COLS])
for r in rows[2:]:
    vals = []
    for full, short in COLS:
        v = r[idx[full]] if full in idx else ""
        if short == "kernel":
            v = v.split("(")[0].replace("void ", "").replace("deno::", "")
        vals.append(v)
    out.writerow(vals)
