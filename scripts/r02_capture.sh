#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -x -q -m gpu -k "captures or real_video" > gpurun_out/r02_pytest_capture.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_pytest_capture.log
timeout 900 python scripts/capture_probe.py --frames 3000 > gpurun_out/r02_capture_probe.log 2> gpurun_out/r02_capture_probe.err; echo "probe rc=$?"; cat gpurun_out/r02_capture_probe.log; tail -3 gpurun_out/r02_capture_probe.err
