#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sift_mixed_gpu.py tests/test_fullsize_gpu.py -q --maxfail=30 -p no:cacheprovider > gpurun_out/r2_4_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2_4_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_4_bench.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_4_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_4_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_4_ncu.log 2>&1
tail -4 gpurun_out/r2_4_tests.log
