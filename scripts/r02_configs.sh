#!/bin/bash
cd "$(dirname "$0")/.."
timeout 1200 python scripts/bench_configs.py --configs 1,2full,4,5,nv12,pcie --videos-per-gpu 4 > gpurun_out/r02_configs_n1.log 2> gpurun_out/r02_configs_n1.err; echo "configs rc=$?"; cut -c1-600 gpurun_out/r02_configs_n1.log; tail -3 gpurun_out/r02_configs_n1.err
