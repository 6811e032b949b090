python scripts/quick_i8.py > gpurun_out/r02_quick_i8_h.jsonl 2> gpurun_out/r02_quick_i8_h.err
python -m pytest tests -m gpu -x -q -k "cfg2 or float64x or device_drift or n_gpus" 2>&1 | tail -5 > gpurun_out/r02_pytest_h.log
tail -n 3 gpurun_out/r02_pytest_h.log
