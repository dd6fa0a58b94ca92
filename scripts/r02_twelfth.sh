#!/bin/bash
cd "$(dirname "$0")/.."
python scripts/decode_case.py > gpurun_out/r02_decode_case_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:jpeg_entropy -c 1 -f -o gpurun_out/r02_jpeg_entropy_v4 python scripts/decode_case.py > gpurun_out/r02_ncu_jpeg.log 2>&1
echo "ncu rc=$?"
python __graft_entry__.py --smoke 2>&1 | tail -2
