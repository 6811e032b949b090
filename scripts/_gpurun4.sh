set -x
python scripts/factor_timing.py > gpurun_out/r02_factor_timing_c.log 2>&1
python scripts/quick_i8.py > gpurun_out/r02_quick_i8_c.jsonl 2> gpurun_out/r02_quick_i8_c.err
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_pytest_c.log
ncu --section SpeedOfLight --section LaunchStats --section Occupancy --section WarpStateStats --section ComputeWorkloadAnalysis --section MemoryWorkloadAnalysis --clock-control none -k regex:'panel_kernel|syrk_kernel|trtri|dual_|pack_kernel' -s 20 -c 36 -o gpurun_out/r02_factor_c python scripts/ncu_r02_drive.py float64 --m=64 > gpurun_out/r02_ncu_factor_c.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_factor_c.csv python scripts/ncu_r02_drive.py float64 --m=64 > /dev/null 2>&1
cat gpurun_out/r02_factor_timing_c.log; cat gpurun_out/r02_quick_i8_c.jsonl | cut -c 1-1500; tail -n 4 gpurun_out/r02_pytest_c.log; ls -la gpurun_out
