#!/bin/bash
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest4.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest4.log
tail -15 gpurun_out/r02_pytest4.log
timeout 900 python scripts/kernel_ab.py > gpurun_out/r02_kernel_ab.log 2> gpurun_out/r02_kernel_ab.err
echo "ab rc=$?"; cat gpurun_out/r02_kernel_ab.log; tail -3 gpurun_out/r02_kernel_ab.err
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_b.log 2> gpurun_out/r02_bench_n1_b.err
echo "bench rc=$?"; tail -c 2500 gpurun_out/r02_bench_n1_b.log; tail -5 gpurun_out/r02_bench_n1_b.err
