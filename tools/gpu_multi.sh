#!/bin/bash
# 2-GPU run: NCCL ring test + default bench line at N=2 (+ reference arm under torchrun)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_m_smi.log 2>&1
timeout 600 python -m pytest tests/test_multigpu.py -q -p no:cacheprovider > gpurun_out/r2_m_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2_m_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_m_bench.log 2>&1
echo "bench rc=$?" >> gpurun_out/r2_m_bench.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 1 --warmup 0 > gpurun_out/r2_m_ref.log 2>&1
echo "ref rc=$?" >> gpurun_out/r2_m_ref.log
tail -3 gpurun_out/r2_m_tests.log; tail -c 600 gpurun_out/r2_m_bench.log; tail -c 300 gpurun_out/r2_m_ref.log
