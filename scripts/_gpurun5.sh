set -x
python scripts/factor_timing.py > gpurun_out/r02_factor_timing_d.log 2>&1
python tests/bench_configs.py cfg5 --frac 0.25 > gpurun_out/r02_cfg5_d.jsonl 2> gpurun_out/r02_cfg5_d.err
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_pytest_d.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_factor_d.csv python scripts/ncu_r02_drive.py float64 --m=64 > /dev/null 2>&1
cat gpurun_out/r02_factor_timing_d.log; cat gpurun_out/r02_cfg5_d.jsonl | cut -c 1-1200; tail -n 3 gpurun_out/r02_cfg5_d.err; tail -n 4 gpurun_out/r02_pytest_d.log
