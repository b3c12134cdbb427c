#!/bin/bash
# usage: gpu_corner.sh <tag> : corner tests + harris workload bench line
tag=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_corners_gpu.py tests/test_edge_cases_gpu.py tests/test_corner_extras_gpu.py -q --maxfail=20 -p no:cacheprovider > gpurun_out/${tag}_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/${tag}_tests.log
timeout 600 python bench.py --workload harris --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench.log 2>&1
tail -5 gpurun_out/${tag}_tests.log; tail -c 1500 gpurun_out/${tag}_bench.log
