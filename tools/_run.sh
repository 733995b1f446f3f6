#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc $?"
tail -8 gpurun_out/pytest_gpu.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r02_n2.json 2> gpurun_out/bench_r02_n2.err
echo "bench n2 rc $?"; tail -c 1500 gpurun_out/bench_r02_n2.json; tail -5 gpurun_out/bench_r02_n2.err
