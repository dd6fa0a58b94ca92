#!/bin/bash
cd "$(dirname "$0")/.."
timeout 300 ncu --set full --import-source on --clock-control none -k regex:decide_kernel -s 12 -c 1 -o gpurun_out/r02_decide_full python scripts/decide_probe.py 1 > gpurun_out/r02_decide_full.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/r02_decide_full.log
ls -la gpurun_out/r02_decide_full.ncu-rep
