#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_2_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_2_ncu.log 2>&1
python tools/diag_outlier.py > gpurun_out/r2_2_diag.log 2>&1
