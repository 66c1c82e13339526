#!/bin/bash
# bench.py under torchrun at N GPUs (the driver's launch line).  Usage: gpurun --gpus N -- bash tools/gpu_scale.sh N tag
N=${1:-2}; tag=${2:-scale}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${tag}_n${N}_gpus.log 2>&1
timeout ${TMO:-300} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/${tag}_n${N}.log 2> gpurun_out/${tag}_n${N}.err
echo "exit $?" >> gpurun_out/${tag}_n${N}.err
tail -c 1500 gpurun_out/${tag}_n${N}.log; tail -5 gpurun_out/${tag}_n${N}.err
