# Round-2 ncu evidence (run under gpurun on ONE B200): --set full captures of the hot kernels, exported to CSV on the box
# (the .ncu-rep files are too large to travel back: gpurun_out is limited to 64 MiB), plus the launch list of bench.py.
set -x
O=gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'solve_kernel_pt|solve_kernel_i8|solve_kernel_tf32' -c 4 -o /tmp/r02_final_solve python scripts/ncu_r02_drive.py float64 float32 float64x float64x4 --m=1000000 > $O/r02_final_solve.log 2>&1
ncu -i /tmp/r02_final_solve.ncu-rep --page raw --csv > $O/r02_final_solve_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:knn_solve_kernel -s 1 -c 1 -o /tmp/r02_final_knn python scripts/ncu_knn_drive.py > $O/r02_final_knn.log 2>&1
ncu -i /tmp/r02_final_knn.ncu-rep --page raw --csv > $O/r02_final_knn_raw.csv 2>/dev/null
ncu -i /tmp/r02_final_knn.ncu-rep --page source --print-source cuda,sass --csv 2>/dev/null | grep -v '^"","",' > $O/r02_final_knn_source_lines.csv
ncu --set full --clock-control none -k regex:'panel_kernel|syrk_kernel|trtri' -s 8 -c 14 -o /tmp/r02_final_factor python scripts/ncu_r02_drive.py float64 --m=64 > $O/r02_final_factor.log 2>&1
ncu -i /tmp/r02_final_factor.ncu-rep --page raw --csv > $O/r02_final_factor_raw.csv 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_final_launches_bench.csv python bench.py --steps 2 --warmup 1 --configs none > $O/r02_final_bench_under_ncu.json 2> $O/r02_final_bench_under_ncu.err
du -sh $O; ls -la $O | tail -12
