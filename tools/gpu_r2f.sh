#!/bin/bash
# Round-2 visit F: TMA wgrad after the quadrant fix (presence diagnostic + suite), TMA analysis bring-up switches.
tag=${1:-r2f}
mkdir -p gpurun_out
timeout 200 python tools/wgrad_diag2.py > gpurun_out/${tag}_wgrad_diag2.log 2>&1
echo "diag2 exit $?" >> gpurun_out/${tag}_wgrad_diag2.log
grep -c "tma" gpurun_out/${tag}_wgrad_diag2.log
python - <<PY
import re
bad = [l for l in open("gpurun_out/${tag}_wgrad_diag2.log") if "old" in l and l.split("old")[1].split("tma")[0].strip() != l.split("tma")[1].strip()]
print("diag2 mismatching lines:", len(bad)); print("".join(bad[:8]))
PY
DENO_WGRAD_TMA=1 timeout 900 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > gpurun_out/${tag}_pytest_wgrad.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest_wgrad.log
tail -6 gpurun_out/${tag}_pytest_wgrad.log
timeout 1200 python tools/an_tma_debug.py > gpurun_out/${tag}_an_tma_debug.log 2>&1
echo "debug exit $?" >> gpurun_out/${tag}_an_tma_debug.log
cat gpurun_out/${tag}_an_tma_debug.log
