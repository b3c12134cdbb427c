#!/bin/bash
# usage: gpu_ncu.sh <tag> <kernel regex> <count> [bench args...]
tag=$1; regex=$2; count=$3; shift 3
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary "$@" > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:${regex}" -c ${count} -o gpurun_out/${tag}_prof python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary "$@" > gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log
