#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_corner_extras_gpu.py tests/test_ingest_gpu.py tests/test_edge_cases_gpu.py tests/test_corners_gpu.py -q --maxfail=30 -p no:cacheprovider > gpurun_out/r2_6_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2_6_tests.log
tail -6 gpurun_out/r2_6_tests.log
