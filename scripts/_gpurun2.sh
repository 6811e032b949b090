set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_pytest_a.log
python scripts/quick_i8.py > gpurun_out/r02_quick_i8_a.jsonl 2> gpurun_out/r02_quick_i8_a.err
ncu --set full --clock-control none -k regex:'solve_kernel_i8|solve_kernel_tf32' -c 3 -o gpurun_out/r02_tc_a python scripts/ncu_r02_drive.py float32 float64x float64x4 > gpurun_out/r02_ncu_tc_a.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_factor_base.csv python scripts/ncu_r02_drive.py float64 --m=64 > /dev/null 2>&1
python scripts/factor_timing.py > gpurun_out/r02_factor_timing_base.log 2>&1
python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err
ls -la gpurun_out
tail -n 5 gpurun_out/r02_pytest_a.log; cat gpurun_out/r02_quick_i8_a.jsonl; tail -n 3 gpurun_out/r02_factor_timing_base.log
