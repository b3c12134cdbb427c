#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sift_mixed_gpu.py tests/test_fullsize_gpu.py tests/test_sift_gpu.py tests/test_edge_cases_gpu.py tests/test_sift_api_gpu.py -q --maxfail=30 -p no:cacheprovider > gpurun_out/r2_3_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2_3_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_3_bench.log 2>&1
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_3_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:gauss_mid_f64|gradient_codes4|descriptor_kernel|orientation_kernel|extrema_repair|extrema_strip" -c 28 -o gpurun_out/r2_3_prof python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_3_ncu.log 2>&1
tail -4 gpurun_out/r2_3_tests.log
