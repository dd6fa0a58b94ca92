#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_edge.py tests/test_gpu_nv12.py tests/test_gpu_hash.py -x -q -s > gpurun_out/r02_pytest7.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest7.log
tail -15 gpurun_out/r02_pytest7.log
timeout 600 python scripts/decode_probe.py --backends native,gpu_hybrid --sessions 1,4,8 --no-cpu > gpurun_out/r02_decode_probe2.log 2> gpurun_out/r02_decode_probe2.err
echo "probe rc=$?"; cat gpurun_out/r02_decode_probe2.log; tail -3 gpurun_out/r02_decode_probe2.err
timeout 600 python scripts/decode_probe.py --backends native --sessions 1,2,4 --batch 256 --frames 512 --no-cpu > gpurun_out/r02_decode_probe3.log 2> gpurun_out/r02_decode_probe3.err
echo "probe3 rc=$?"; cat gpurun_out/r02_decode_probe3.log; tail -3 gpurun_out/r02_decode_probe3.err
