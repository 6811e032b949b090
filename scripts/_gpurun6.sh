python scripts/debug_factor.py > gpurun_out/r02_debug_factor.log 2>&1
cat gpurun_out/r02_debug_factor.log
