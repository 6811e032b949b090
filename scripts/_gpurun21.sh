python scripts/tile_timing.py > gpurun_out/r02_tile_timing2.log 2>&1
python -m pytest tests -m gpu -x -q -k "tile_width or full_size_properties_cfg2 or staged or shard or exact_hits or cfg1" 2>&1 | tail -4 > gpurun_out/r02_pytest_n.log
cat gpurun_out/r02_tile_timing2.log; tail -n 3 gpurun_out/r02_pytest_n.log
