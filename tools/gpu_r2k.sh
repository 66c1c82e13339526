#!/bin/bash
# Round-2 visit K (2 GPUs): DP parity tests (all collective flavours incl. the fused multicast kernel), N=2 bench fused vs NCCL.
tag=${1:-r2k}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${tag}_gpus.log 2>&1
timeout 600 python -m pytest tests/test_gpu_ddp.py -q -m gpu -p no:cacheprovider > gpurun_out/${tag}_pytest_ddp.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest_ddp.log
grep -E "^FAILED|passed|failed|exit|Error|error" gpurun_out/${tag}_pytest_ddp.log | cut -c1-300 | tail -25
for f in 1 0; do
  DENO_DP_FUSED=$f timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$f \
    bench.py --gpus 2 --steps 200 --warmup 5 > gpurun_out/${tag}_n2_fused$f.log 2> gpurun_out/${tag}_n2_fused$f.err
  echo "bench fused=$f exit $?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_n2_fused$f.log").read().strip().splitlines()[-1])
    print(" value %.0f e2e %.0f ms %.3f allreduce: %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["config"].get("allreduce")))
except Exception as e:
    print(" parse failed", e)
PY
  tail -3 gpurun_out/${tag}_n2_fused$f.err
done
timeout 300 python bench.py --steps 200 --no-secondary --cpu-steps 1 > gpurun_out/${tag}_n1.log 2>&1
python - <<PY
import json
d=json.loads(open("gpurun_out/${tag}_n1.log").read().strip().splitlines()[-1])
print("N=1 value %.0f ms %.3f" % (d["value"], d["ms_per_step"]))
PY
