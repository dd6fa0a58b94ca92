#!/bin/bash
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8; nproc
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_n8.log 2> gpurun_out/r02_bench_n8.err
echo "bench n8 rc=$?"; tail -c 1500 gpurun_out/r02_bench_n8.log; tail -5 gpurun_out/r02_bench_n8.err
timeout 300 python -m pytest tests/test_gpu_multi.py -x -q -k "sharded or library or service" > gpurun_out/r02_pytest_n8.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_n8.log
