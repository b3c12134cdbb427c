#!/bin/bash
# Round-1 ncu evidence, run on the GPU box (gpurun -- 'bash tools/capture_profiles.sh'): for each workload the bench
# first runs WITHOUT ncu (must exit 0), then the launch list (gpu__time_duration per launch) and one --set full
# capture of the named kernels.  Outputs land in gpurun_out/; profiles/summarize.py turns them into profiles/r01_*.
set -u
out=gpurun_out
mkdir -p $out
run() {  # name, workload flags, kernel regex, capture count
  local name=$1 flags=$2 regex=$3 count=$4
  python bench.py $flags --steps 2 --warmup 1 --no-cpu-baseline > $out/plain_$name.log 2>&1 || { echo "plain $name failed"; return; }
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/r01_launches_$name.csv \
      python bench.py $flags --steps 1 --warmup 1 --no-cpu-baseline > $out/ncu_l_$name.log 2>&1
  ncu --set full --clock-control none --import-source on -k "regex:$regex" -c $count -f -o $out/r01_${name}_kernels \
      python bench.py $flags --steps 1 --warmup 1 --no-cpu-baseline > $out/ncu_f_$name.log 2>&1
}
only=${1:-all}
want() { [ "$only" = all ] || [ "$only" = "$1" ]; }
want sift && run sift "--workload sift --batch 8" "gauss_octave_packed|extrema_strip|gradient_field4|orientation_kernel|descriptor_kernel|upsample2_block" 40
want harris && run harris "--workload harris" "harris5_f32|lambda5_ring|threshold_mask" 3
want match && run match "--workload match" "match_tc2|match_rerank|match_prep" 3
ls -la $out | tail -20
