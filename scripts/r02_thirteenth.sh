#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_i420.py -x -q -m gpu > gpurun_out/r02_pytest13.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest13.log
ESD_DEC_TIMING=2 timeout 300 python scripts/decode_trace.py --grid-cap 24 --passes 2 > gpurun_out/r02_timeline.log 2> gpurun_out/r02_timeline.err; echo "trace rc=$?"; cat gpurun_out/r02_timeline.log
grep -c timeline gpurun_out/r02_timeline.err
