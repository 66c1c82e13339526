#!/bin/bash
# bench.py under torchrun at N GPUs (the driver's launch line), default exchange (fused multicast) only.
#   gpurun --gpus N -- bash tools/gpu_scale_fused.sh N tag
N=${1:-8}; tag=${2:-scale}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${tag}_n${N}_gpus.log 2>&1
NCCL_DEBUG=WARN timeout ${TMO:-400} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 200 --warmup 5 > gpurun_out/${tag}_n${N}.log 2> gpurun_out/${tag}_n${N}.err
echo "N=$N exit $?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_n${N}.log").read().strip().splitlines()[-1])
    print(" value %.0f e2e %.0f ms %.4f allreduce: %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["config"].get("allreduce")[:60]))
    for k, w in (d.get("workloads") or {}).items():
        print("  ", k, {kk: w.get(kk) for kk in ("value", "ms_per_step", "n_gpus", "error")})
except Exception as e:
    print(" parse failed", e)
PY
grep -v "OMP_NUM_THREADS\|^\*\*\*\*" gpurun_out/${tag}_n${N}.err | tail -3
