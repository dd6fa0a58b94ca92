#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_i420.py tests/test_gpu_nv12.py -x -q -m gpu > gpurun_out/r02_pytest11.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_pytest11.log
for v in "--grid-cap 24" "--grid-cap 0" "--grid-cap 24 --sync-each-pass 0" "--grid-cap 8" "--sessions 12 --grid-cap 24"; do
timeout 300 python scripts/decode_trace.py $v > gpurun_out/r02_trace.log 2> gpurun_out/r02_trace.err; echo "trace [$v] rc=$?"; cat gpurun_out/r02_trace.log; tail -2 gpurun_out/r02_trace.err
done
