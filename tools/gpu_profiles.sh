#!/bin/bash
# evidence for profiles/: launch lists of the three workloads + full captures of the dominant kernels
mkdir -p gpurun_out
run() { tag=$1; shift; python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary "$@" > gpurun_out/${tag}_plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary "$@" > gpurun_out/${tag}_ncu.log 2>&1; }
run r02l_sift
run r02l_harris --workload harris
run r02l_match --workload match
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r02f_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:gauss_octave_packed|gauss_mid_f64|extrema_strip|extrema_repair|gradient_codes4|orientation_kernel|orientation_recheck|descriptor_coded" -c 40 -o gpurun_out/r02f_sift python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r02f_ncu.log 2>&1
ls -la gpurun_out/r02*
