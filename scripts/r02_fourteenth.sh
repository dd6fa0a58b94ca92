#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_i420.py -x -q -m gpu > gpurun_out/r02_pytest14.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest14.log
for mc in 8 32; do
CUDA_DEVICE_MAX_CONNECTIONS=$mc timeout 300 python scripts/decode_trace.py --grid-cap 24 --passes 3 > gpurun_out/r02_mc$mc.log 2> gpurun_out/r02_mc$mc.err; echo "trace mc=$mc rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02_mc$mc.log')); print(d['frames_per_s'], d['seconds'])"
done
CUDA_DEVICE_MAX_CONNECTIONS=32 ESD_DEC_TIMING=2 timeout 300 python scripts/decode_trace.py --grid-cap 24 --passes 2 > gpurun_out/r02_timeline32.log 2> gpurun_out/r02_timeline32.err; echo "trace rc=$?"; cat gpurun_out/r02_timeline32.log | cut -c1-120
