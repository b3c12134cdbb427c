#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_ap_topo.log 2>&1
NCCL_DEBUG=INFO timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload allpairs --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ap_bench.log 2>&1
echo "rc=$?" >> gpurun_out/r2_ap_bench.log
grep -i "via\|P2P\|NVLS\|SHM\|NET/" gpurun_out/r2_ap_bench.log | head -12; cat gpurun_out/r2_ap_topo.log | head -8; tail -c 900 gpurun_out/r2_ap_bench.log
