#!/usr/bin/env python3
"""Summarise an `ncu --page source --csv` export: opcode mix, stall reasons, hottest SASS lines.
   ncu -i rep.ncu-rep --page source --csv --kernel-name K --launch-count 1 > k.csv; python tools/ncu_source_summary.py k.csv"""
import collections
import csv
import sys


def main(path, ntop=30):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr, data = rows[hi], [r for r in rows[hi + 1:] if len(r) == len(rows[hi]) and r[0].startswith("0x")]
    iS, iE, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    ops, samp, tot = collections.Counter(), collections.Counter(), 0
    for r in data:
        toks = r[iS].split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        op = op.split(".")[0]
        n = int(r[iE]); ops[op] += n; tot += n; samp[op] += int(r[iN])
    print("kernel:", rows[0][1] if rows[0] else "?", " warp-instructions:", tot, " samples:", sum(samp.values()))
    for op, n in ops.most_common(22):
        print(f"  {op:10s} {n:11d} {100 * n / tot:5.1f}%  samples {samp[op]}")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    agg = collections.Counter()
    for r in data:
        for i in stall_cols:
            agg[hdr[i]] += int(r[i])
    print("stalls:", [(k, v) for k, v in agg.most_common(8)])
    for r in sorted(data, key=lambda r: -int(r[iN]))[:ntop]:
        print(f"  {r[iN]:>6s} smp {r[iE]:>9s} exe  {r[iS][:90]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
