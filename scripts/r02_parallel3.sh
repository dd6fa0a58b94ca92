#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_decode.py -x -q -m gpu > gpurun_out/r02_pytest_par3.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_par3.log
ESD_DEC_TIMING=1 timeout 600 python scripts/decode_probe.py --backends native --sessions 1,2,4 --batch 64 --frames 1024 --no-cpu > gpurun_out/r02_probe_par3.log 2> gpurun_out/r02_probe_par3.err
echo "probe rc=$?"; cat gpurun_out/r02_probe_par3.log; grep timing gpurun_out/r02_probe_par3.err | tail -2
ESD_DEC_STAGE_THREADS=1 ESD_DEC_TIMING=1 timeout 600 python scripts/decode_probe.py --backends native --sessions 1 --batch 64 --frames 1024 --no-cpu > gpurun_out/r02_probe_par3b.log 2> gpurun_out/r02_probe_par3b.err
echo "probe (1 staging thread) rc=$?"; cat gpurun_out/r02_probe_par3b.log; grep timing gpurun_out/r02_probe_par3b.err | tail -1
ESD_DEC_TIMING=1 timeout 600 python scripts/decode_probe.py --backends native --sessions 1,4 --batch 256 --frames 1024 --no-cpu > gpurun_out/r02_probe_par3c.log 2> gpurun_out/r02_probe_par3c.err
echo "probe256 rc=$?"; cat gpurun_out/r02_probe_par3c.log; grep timing gpurun_out/r02_probe_par3c.err | tail -1
