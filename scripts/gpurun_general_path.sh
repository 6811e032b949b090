set -x
mkdir -p gpurun_out
timeout 240 python scripts/gj_timing.py > gpurun_out/r02_gj_timing.jsonl 2> gpurun_out/r02_gj_timing.err
cat gpurun_out/r02_gj_timing.jsonl; tail -5 gpurun_out/r02_gj_timing.err
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gj.log 2>&1
tail -8 gpurun_out/r02_pytest_gj.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_gj.csv python scripts/gj_ncu_drive.py > gpurun_out/r02_gj_ncu.log 2>&1
tail -3 gpurun_out/r02_gj_ncu.log
