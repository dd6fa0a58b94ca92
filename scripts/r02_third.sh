#!/bin/bash
# round 2, third GPU call: GPU decode tests + throughput probe, decide pass after the kernel rework
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_decode.py -x -q -s > gpurun_out/r02_pytest3.log 2>&1
echo "pytest decode rc=$?" >> gpurun_out/r02_pytest3.log
tail -30 gpurun_out/r02_pytest3.log
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py tests/test_gpu_plugin.py -x -q -s -k "decide or filter or plugin or golden" > gpurun_out/r02_pytest3b.log 2>&1
echo "pytest decide rc=$?" >> gpurun_out/r02_pytest3b.log
tail -8 gpurun_out/r02_pytest3b.log
timeout 900 python scripts/decode_probe.py > gpurun_out/r02_decode_probe.log 2> gpurun_out/r02_decode_probe.err
echo "probe rc=$?"; cat gpurun_out/r02_decode_probe.log; tail -5 gpurun_out/r02_decode_probe.err
