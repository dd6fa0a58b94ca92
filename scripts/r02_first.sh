#!/bin/bash
# round 2, first GPU call: what the box offers for decode, then the GPU suite on the refactored library
bash scripts/probe_box.sh
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest1.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest1.log
tail -5 gpurun_out/r02_pytest1.log
tail -60 gpurun_out/probe_box.log
