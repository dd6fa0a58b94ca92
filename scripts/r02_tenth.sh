#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_i420.py tests/test_gpu_nv12.py tests/test_gpu_decode.py -x -q -m gpu > gpurun_out/r02_pytest10.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_pytest10.log
for s in 1 4 8; do
ESD_DEC_TIMING=1 timeout 600 python scripts/decode_probe.py --backends native --sessions $s --batch 256 --frames 768 --no-cpu > gpurun_out/r02_probe_t$s.log 2> gpurun_out/r02_probe_t$s.err
echo "probe s=$s rc=$?"; cat gpurun_out/r02_probe_t$s.log; grep "esd_decode timing" gpurun_out/r02_probe_t$s.err | tail -6
done
