#!/bin/bash
# usage: gpu_ui.sh <tag> : UI-range tests + timing of a few non-default parameter points
tag=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sift_mixed_gpu.py -q --maxfail=20 -p no:cacheprovider > gpurun_out/${tag}_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/${tag}_tests.log
: > gpurun_out/${tag}_ui.log
for pt in "1.6 2" "2.5 3" "3.0 5" "3.0 2" "2.0 4 8"; do python tools/ui_point.py $pt >> gpurun_out/${tag}_ui.log 2>&1; done
tail -5 gpurun_out/${tag}_tests.log; cat gpurun_out/${tag}_ui.log
