#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_i420.py -x -q -m gpu > gpurun_out/r02_pytest12.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_pytest12.log
for v in "--grid-cap 24" "--sessions 4 --grid-cap 24" "--sessions 16 --grid-cap 24 --batch 128"; do
ESD_DEC_TIMING=1 timeout 300 python scripts/decode_trace.py $v > gpurun_out/r02_trace.log 2> gpurun_out/r02_trace.err; echo "trace [$v] rc=$?"; cat gpurun_out/r02_trace.log; grep timing gpurun_out/r02_trace.err | head -4
done
ESD_DEC_TIMING=1 timeout 600 python scripts/decode_probe.py --backends native --sessions 12,16 --batch 256 --frames 768 --no-cpu > gpurun_out/r02_probe_t16.log 2> gpurun_out/r02_probe_t16.err
cat gpurun_out/r02_probe_t16.log; grep timing gpurun_out/r02_probe_t16.err | tail -3
