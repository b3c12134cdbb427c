python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/final_default.log 2>&1; tail -1 gpurun_out/final_default.log | cut -c1-300
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/final_ref.log 2>&1; tail -1 gpurun_out/final_ref.log | cut -c1-200
python bench.py --workload harris --no-cpu-baseline > gpurun_out/final_harris.log 2>&1; tail -1 gpurun_out/final_harris.log | cut -c1-200
python bench.py --workload match --no-cpu-baseline > gpurun_out/final_match.log 2>&1; tail -1 gpurun_out/final_match.log | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
