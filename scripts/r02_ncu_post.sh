#!/bin/bash
cd "$(dirname "$0")/.."
timeout 300 python scripts/decode_probe.py --backends native --sessions 1 --batch 64 --frames 128 --no-cpu > gpurun_out/r02_ncu_post_plain.log 2>&1; echo "plain rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"jpeg_color_kernel|jpeg_idct_sparse_kernel" -s 4 -c 2 -o gpurun_out/r02_post_full python scripts/decode_probe.py --backends native --sessions 1 --batch 64 --frames 128 --no-cpu > gpurun_out/r02_ncu_post.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/r02_ncu_post.log
