nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err
echo "rc=$?"; tail -n 4 gpurun_out/r02_bench_n8.err; head -c 400 gpurun_out/r02_bench_n8.json
