#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_decode.py -x -q -m gpu > gpurun_out/r02_pytest16.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest16.log
for v in "--sessions 8" "--sessions 4" "--sessions 6"; do
timeout 300 python scripts/decode_trace.py --grid-cap 24 --passes 3 --frames 1024 $v > gpurun_out/r02_lanes.log 2> gpurun_out/r02_lanes.err; echo "trace [$v] rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02_lanes.log')); print(d['frames_per_s'], d['seconds'])"; tail -2 gpurun_out/r02_lanes.err
done
ESD_DEC_LANES=1 timeout 300 python scripts/decode_trace.py --grid-cap 24 --passes 3 --frames 1024 --sessions 8 > gpurun_out/r02_lanes1.log 2> gpurun_out/r02_lanes1.err; python -c "
import json; d=json.load(open('gpurun_out/r02_lanes1.log')); print('one lane', d['frames_per_s'], d['seconds'])"
ESD_DEC_TIMING=1 timeout 600 python scripts/decode_probe.py --backends native --sessions 4,8 --batch 256 --frames 1024 --no-cpu > gpurun_out/r02_probe_lanes.log 2> gpurun_out/r02_probe_lanes.err
cat gpurun_out/r02_probe_lanes.log; grep timing gpurun_out/r02_probe_lanes.err | tail -3
timeout 200 python scripts/decide_probe.py > gpurun_out/r02_decide_probe.log 2>&1; cat gpurun_out/r02_decide_probe.log
timeout 300 ncu --metrics gpu__time_duration.sum -k regex:decide_kernel -c 12 --csv --log-file gpurun_out/r02_decide_ncu.csv python scripts/decide_probe.py 2 > /dev/null 2>&1; tail -13 gpurun_out/r02_decide_ncu.csv | cut -d, -f5,12-15
