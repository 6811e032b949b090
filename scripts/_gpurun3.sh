set -x
python scripts/quick_i8.py > gpurun_out/r02_quick_i8_b.jsonl 2> gpurun_out/r02_quick_i8_b.err
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_pytest_b.log
ncu --set full --clock-control none -k regex:'solve_kernel_i8|solve_kernel_tf32|solve_kernel_pt' -c 4 -o gpurun_out/r02_tc_b python scripts/ncu_r02_drive.py float64 float32 float64x float64x4 > gpurun_out/r02_ncu_tc_b.log 2>&1
cat gpurun_out/r02_quick_i8_b.jsonl; tail -n 4 gpurun_out/r02_pytest_b.log; tail -n 3 gpurun_out/r02_quick_i8_b.err
