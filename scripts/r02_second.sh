#!/bin/bash
# round 2, second GPU call: new GPU tests, the bench as the driver runs it (both arms), NVDEC capabilities
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_bench.py -x -q -s > gpurun_out/r02_pytest2.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest2.log
tail -25 gpurun_out/r02_pytest2.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.log 2> gpurun_out/r02_bench_n1.err
echo "bench rc=$?"; tail -c 3000 gpurun_out/r02_bench_n1.log; tail -5 gpurun_out/r02_bench_n1.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_ref.log 2> gpurun_out/r02_bench_ref.err
echo "ref rc=$?"; tail -c 1500 gpurun_out/r02_bench_ref.log
timeout 60 scripts/probes/cuvid_caps > gpurun_out/cuvid_caps.log 2>&1; echo "caps rc=$?"; cat gpurun_out/cuvid_caps.log
