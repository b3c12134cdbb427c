#!/bin/bash
# usage: gpu_launches.sh <tag> [bench args...]   -- launch list (one ncu pass) of one workload
tag=$1; shift
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary "$@" > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary "$@" > gpurun_out/${tag}_ncu.log 2>&1
tail -c 600 gpurun_out/${tag}_plain.log; wc -l gpurun_out/${tag}_launches.csv
