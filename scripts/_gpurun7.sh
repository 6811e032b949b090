python scripts/factor_timing.py > gpurun_out/r02_factor_timing_e.log 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_pytest_e.log
python scripts/ncu_knn_drive.py 2000000 > gpurun_out/r02_knn_e.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:knn_solve_kernel -s 1 -c 1 -o gpurun_out/r02_knn_e python scripts/ncu_knn_drive.py > gpurun_out/r02_ncu_knn_e.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_factor_e.csv python scripts/ncu_r02_drive.py float64 --m=64 > /dev/null 2>&1
cat gpurun_out/r02_factor_timing_e.log; tail -n 4 gpurun_out/r02_pytest_e.log; cat gpurun_out/r02_knn_e.log | tail -2; ls -la gpurun_out | tail -8
