#!/bin/bash
# round 2, GPU call 1: correctness of the certified mode + regression + first timings
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_1_smi.log 2>&1
timeout 900 python -m pytest tests/test_sift_mixed_gpu.py -q --maxfail=30 -p no:cacheprovider > gpurun_out/r2_1_mixed.log 2>&1
echo "mixed rc=$?" >> gpurun_out/r2_1_mixed.log
timeout 900 python -m pytest tests/test_fullsize_gpu.py -q --maxfail=30 -p no:cacheprovider > gpurun_out/r2_1_full.log 2>&1
echo "full rc=$?" >> gpurun_out/r2_1_full.log
timeout 1200 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider --deselect tests/test_sift_mixed_gpu.py --deselect tests/test_fullsize_gpu.py > gpurun_out/r2_1_rest.log 2>&1
echo "rest rc=$?" >> gpurun_out/r2_1_rest.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2_1_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r2_1_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_1_bench_mixed.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sift-dtype f32 > gpurun_out/r2_1_bench_f32.log 2>&1
tail -3 gpurun_out/r2_1_mixed.log gpurun_out/r2_1_full.log gpurun_out/r2_1_rest.log gpurun_out/r2_1_smoke.log
tail -c 1500 gpurun_out/r2_1_bench_mixed.log
