#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_nv12.py tests/test_gpu_i420.py -x -q -m gpu > gpurun_out/r02_pytest20.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest20.log
timeout 300 python scripts/kernel_ab.py --cases nv12,720p,1080p --strides 0 --seconds 1.5 2>&1 | tail -4
timeout 300 python scripts/bench_configs.py --configs nv12 2>&1 | tail -1 | cut -c1-400
