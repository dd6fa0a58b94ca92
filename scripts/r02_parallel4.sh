#!/bin/bash
cd "$(dirname "$0")/.."
for v in "--sessions 4 --batch 64" "--sessions 4 --batch 128" "--sessions 4 --batch 256" "--sessions 3 --batch 256" "--sessions 2 --batch 256" "--sessions 1 --batch 256" "--sessions 1 --batch 64"; do
timeout 300 python scripts/decode_trace.py --grid-cap 24 --passes 4 --frames 1024 $v > gpurun_out/r02_par_sweep.log 2> gpurun_out/r02_par_sweep.err; echo -n "trace [$v] rc=$? "; python -c "
import json; d=json.load(open('gpurun_out/r02_par_sweep.log')); print(round(d['frames_per_s']), d['seconds'])"
done
