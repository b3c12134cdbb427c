#!/bin/bash
tag=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sift_gpu.py tests/test_sift_mixed_gpu.py tests/test_sift_api_gpu.py tests/test_edge_cases_gpu.py -q --maxfail=20 -p no:cacheprovider > gpurun_out/${tag}_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/${tag}_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary --sift-dtype f64 --batch 8 > gpurun_out/${tag}_bench.log 2>&1
tail -3 gpurun_out/${tag}_tests.log
