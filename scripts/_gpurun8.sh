python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_pytest_f.log
python scripts/ncu_knn_drive.py 4000000 > gpurun_out/r02_knn_f.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:knn_solve_kernel -s 1 -c 1 -o gpurun_out/r02_knn_f python scripts/ncu_knn_drive.py > gpurun_out/r02_ncu_knn_f.log 2>&1
tail -n 4 gpurun_out/r02_pytest_f.log; cat gpurun_out/r02_knn_f.log | tail -2
