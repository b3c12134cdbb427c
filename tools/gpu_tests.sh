#!/bin/bash
tag=$1; shift
mkdir -p gpurun_out
timeout 1200 python -m pytest "$@" -q --maxfail=30 -p no:cacheprovider > gpurun_out/${tag}_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/${tag}_tests.log
grep -n "^E  \|^FAILED\|passed\|failed" gpurun_out/${tag}_tests.log | head -40
