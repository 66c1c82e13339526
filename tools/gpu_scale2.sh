#!/bin/bash
# bench.py under torchrun at N GPUs (the driver's launch line), fused multicast exchange (default) and NCCL.  Usage:
#   gpurun --gpus N -- bash tools/gpu_scale2.sh N tag
N=${1:-2}; tag=${2:-scale}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${tag}_n${N}_gpus.log 2>&1
for f in 1 0; do
  DENO_DP_FUSED=$f NCCL_DEBUG=WARN timeout ${TMO:-400} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$f \
      bench.py --gpus $N --steps 200 --warmup 5 $EXTRA > gpurun_out/${tag}_n${N}_fused$f.log 2> gpurun_out/${tag}_n${N}_fused$f.err
  echo "N=$N fused=$f exit $?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_n${N}_fused$f.log").read().strip().splitlines()[-1])
    print(" value %.0f e2e %.0f ms %.4f allreduce: %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["config"].get("allreduce")[:60]))
    for k, w in (d.get("workloads") or {}).items():
        print("  ", k, {kk: w.get(kk) for kk in ("value", "ms_per_step", "n_gpus", "error")})
except Exception as e:
    print(" parse failed", e)
PY
  grep -v "OMP_NUM_THREADS\|^\*\*\*\*" gpurun_out/${tag}_n${N}_fused$f.err | tail -3
done
