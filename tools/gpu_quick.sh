#!/bin/bash
# usage: gpu_quick.sh <tag> [pytest targets...]: selected tests + main bench line only
tag=$1; shift
mkdir -p gpurun_out
timeout 900 python -m pytest "$@" -q --maxfail=20 -p no:cacheprovider > gpurun_out/${tag}_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/${tag}_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/${tag}_bench.log 2>&1
tail -3 gpurun_out/${tag}_tests.log
