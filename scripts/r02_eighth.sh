#!/bin/bash
cd "$(dirname "$0")/.."
python scripts/decode_case.py > gpurun_out/r02_decode_case_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:jpeg_entropy -c 1 -o gpurun_out/r02_jpeg_entropy python scripts/decode_case.py > gpurun_out/r02_ncu_jpeg.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r02_ncu_jpeg.log
python scripts/decode_case.py > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_jpeg_launches.csv python scripts/decode_case.py > /dev/null 2>&1
echo "launches rc=$?"; grep -v "^==" gpurun_out/r02_jpeg_launches.csv | awk -F'","' '{print $5, $NF}' | tail -12
