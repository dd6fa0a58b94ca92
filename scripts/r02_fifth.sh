#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_cube.py tests/test_gpu_decode.py tests/test_gpu_nv12.py -x -q -s > gpurun_out/r02_pytest5.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest5.log
tail -12 gpurun_out/r02_pytest5.log
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --no-compressed --e2e-modes gather,pageable_gather > gpurun_out/r02_bench_pool.log 2>&1; echo "bench(pool) rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_pool.log') if l.startswith('{')][-1])
print("value", d["value"], "frac", d["roofline"]["frac"], {k:v["value"] for k,v in d["e2e"]["modes"].items()})
PY
python scripts/kernel_ab.py --cases 720p --strides 0 --seconds 0.05 > gpurun_out/r02_plain_720p.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fused_score -s 6 -c 2 -o gpurun_out/r02_fused_720p python scripts/kernel_ab.py --cases 720p --strides 0 --seconds 0.05 > gpurun_out/r02_ncu_720p.log 2>&1
echo "ncu 720p rc=$?"
python scripts/kernel_ab.py --cases 1080p --strides 0 --seconds 0.05 > gpurun_out/r02_plain_1080p.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fused_score -s 6 -c 2 -o gpurun_out/r02_fused_1080p python scripts/kernel_ab.py --cases 1080p --strides 0 --seconds 0.05 > gpurun_out/r02_ncu_1080p.log 2>&1
echo "ncu 1080p rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -3
