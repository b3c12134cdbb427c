#!/bin/bash
# usage: gpu_scale.sh <tag> <n>: default bench line under torchrun at N GPUs
tag=$1; n=$2
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/${tag}_bench.log 2>&1
echo "bench rc=$?" >> gpurun_out/${tag}_bench.log
tail -c 400 gpurun_out/${tag}_bench.log
