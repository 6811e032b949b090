nvidia-smi -L
python -m pytest tests/test_multigpu_gpu.py -x -q 2>&1 | tail -25 > gpurun_out/r02_pytest_mgpu.log
tail -n 12 gpurun_out/r02_pytest_mgpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err
tail -n 5 gpurun_out/r02_bench_n2.err; head -c 600 gpurun_out/r02_bench_n2.json
