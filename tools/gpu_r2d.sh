#!/bin/bash
# Round-2 visit D/E: full GPU suite with the default kernels, smoke, full bench, TMA wgrad batch-presence diagnostic.
tag=${1:-r2d}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -15 gpurun_out/${tag}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/${tag}_smoke.log
tail -3 gpurun_out/${tag}_smoke.log
SECONDS=0
timeout 900 python bench.py > gpurun_out/${tag}_bench.log 2> gpurun_out/${tag}_bench.err
echo "bench exit $? after ${SECONDS}s" | tee -a gpurun_out/${tag}_bench.err
tail -c 6000 gpurun_out/${tag}_bench.log
timeout 200 python tools/wgrad_diag2.py > gpurun_out/${tag}_wgrad_diag2.log 2>&1
echo "diag2 exit $?" >> gpurun_out/${tag}_wgrad_diag2.log
cat gpurun_out/${tag}_wgrad_diag2.log
