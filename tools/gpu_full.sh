#!/bin/bash
# whole GPU test suite + smoke + default bench; logs under gpurun_out/<tag>_*
tag=${1:-full}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/${tag}_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/${tag}_tests.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/${tag}_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/${tag}_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.log 2>&1
tail -4 gpurun_out/${tag}_tests.log; tail -2 gpurun_out/${tag}_smoke.log
