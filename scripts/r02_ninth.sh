#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_decode.py -x -q > gpurun_out/r02_pytest9.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest9.log
timeout 600 python scripts/decode_probe.py --backends native --sessions 1,4,8 --batch 128 --frames 256 --no-cpu > gpurun_out/r02_decode_probe4.log 2> gpurun_out/r02_decode_probe4.err
echo "probe4 rc=$?"; cat gpurun_out/r02_decode_probe4.log; tail -3 gpurun_out/r02_decode_probe4.err
timeout 600 python scripts/decode_probe.py --backends native --sessions 4,8 --batch 256 --frames 512 --no-cpu > gpurun_out/r02_decode_probe5.log 2> gpurun_out/r02_decode_probe5.err
echo "probe5 rc=$?"; cat gpurun_out/r02_decode_probe5.log; tail -3 gpurun_out/r02_decode_probe5.err
python scripts/decode_case.py > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum --clock-control none -k regex:jpeg_entropy -c 1 python scripts/decode_case.py 2>&1 | grep -E "duration|ratio|inst_executed"
