#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_decode.py -x -q -m gpu > gpurun_out/r02_pytest_idct.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02_pytest_idct.log
for v in "--sessions 4 --batch 64" "--sessions 1 --batch 256"; do
ESD_DEC_TIMING=1 timeout 300 python scripts/decode_trace.py --grid-cap 24 --passes 4 --frames 1024 $v > gpurun_out/r02_par_sweep.log 2> gpurun_out/r02_par_sweep.err; echo -n "trace [$v] rc=$? "; python -c "
import json; d=json.load(open('gpurun_out/r02_par_sweep.log')); print(round(d['frames_per_s']), d['seconds'])"; grep timing gpurun_out/r02_par_sweep.err | tail -1
done
