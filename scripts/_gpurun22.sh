python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02_pytest_final.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_final.log 2>&1
tail -n 3 gpurun_out/r02_pytest_final.log; tail -n 2 gpurun_out/r02_bench_final.err; head -c 300 gpurun_out/r02_bench_final.json; tail -n 1 gpurun_out/r02_smoke_final.log
