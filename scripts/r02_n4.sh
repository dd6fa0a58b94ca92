#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02_bench_n4.log 2> gpurun_out/r02_bench_n4.err
echo "bench n4 rc=$?"; tail -c 600 gpurun_out/r02_bench_n4.log; tail -3 gpurun_out/r02_bench_n4.err
