#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_decode.py -x -q -m gpu > gpurun_out/r02_pytest_par.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_par.log
ESD_DEC_TIMING=1 timeout 600 python scripts/decode_probe.py --backends native --sessions 1,2,4,8 --batch 256 --frames 1024 --no-cpu > gpurun_out/r02_probe_par.log 2> gpurun_out/r02_probe_par.err
echo "probe rc=$?"; cat gpurun_out/r02_probe_par.log; grep timing gpurun_out/r02_probe_par.err | tail -3
ESD_DEC_TIMING=1 timeout 600 python scripts/decode_probe.py --backends native --sessions 1,4 --batch 64 --frames 1024 --no-cpu > gpurun_out/r02_probe_par64.log 2> gpurun_out/r02_probe_par64.err
echo "probe64 rc=$?"; cat gpurun_out/r02_probe_par64.log; grep timing gpurun_out/r02_probe_par64.err | tail -2
