#!/usr/bin/env python3
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: finds one training step (the launches
between two consecutive adam_flat_kernel launches) and prints per-kernel totals and shares.
python tools/launchlist_summary.py gpurun_out/ll_launches.csv"""
import collections
import csv
import sys

rows = []
with open(sys.argv[1]) as f:
    lines = [ln for ln in f if ln.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    val = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = val / 1e3 if unit in ("ns", "nsecond") else (val if unit in ("us", "usecond") else val * 1e3)
    rows.append((r["Kernel Name"], us))
adam = [i for i, (k, _) in enumerate(rows) if "adam_flat" in k]
if len(adam) < 2:
    print(f"{len(rows)} launches, fewer than two adam_flat launches in the window")
    sys.exit(1)
step = rows[adam[0] + 1:adam[1] + 1]
tot = sum(us for _, us in step)
agg = collections.OrderedDict()
for k, us in step:
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += us
ours = sum(v[1] for k, v in agg.items() if "deno::" in k or "tc::" in k)
print(f"# one training step (between two adam_flat launches): {len(step)} launches, {tot:.1f} us total "
      f"(cold-cache, serialised: compare SHARES); kernels of libdeno_spectral.so: {100 * ours / tot:.1f}%")
for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{us:9.1f} us {100 * us / tot:5.1f}% x{c:3d}  {k[:110]}")
