#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ingest_gpu.py tests/test_sift_api_gpu.py tests/test_fullsize_gpu.py -q --maxfail=30 -p no:cacheprovider > gpurun_out/r2_5_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2_5_tests.log
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2_5_bench.log 2>&1
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r2_5_ref.log 2>&1
tail -4 gpurun_out/r2_5_tests.log; tail -5 gpurun_out/r2_5_bench.log | cut -c1-600
