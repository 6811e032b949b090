python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_pytest_g.log
python scripts/ncu_knn_drive.py 4000000 > gpurun_out/r02_knn_g.log 2>&1
python scripts/quick_i8.py > gpurun_out/r02_quick_i8_g.jsonl 2> gpurun_out/r02_quick_i8_g.err
python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_g.json 2> gpurun_out/r02_bench_g.err
tail -n 4 gpurun_out/r02_pytest_g.log; tail -n 1 gpurun_out/r02_knn_g.log; tail -n 2 gpurun_out/r02_bench_g.err; head -c 300 gpurun_out/r02_bench_g.json
