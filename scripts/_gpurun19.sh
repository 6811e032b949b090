python scripts/tile_timing.py > gpurun_out/r02_tile_timing.log 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02_pytest_l.log
cat gpurun_out/r02_tile_timing.log; tail -n 3 gpurun_out/r02_pytest_l.log
