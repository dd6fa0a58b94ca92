#!/bin/bash
# two GPUs: the multi-device product surface on distinct devices, the bench under torchrun (config 3 on the ranks)
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_decode.py tests/test_gpu_hash.py -x -q -s > gpurun_out/r02_pytest_n2.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_n2.log
tail -8 gpurun_out/r02_pytest_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_n2.log 2> gpurun_out/r02_bench_n2.err
echo "bench n2 rc=$?"; tail -c 3500 gpurun_out/r02_bench_n2.log; tail -5 gpurun_out/r02_bench_n2.err
timeout 300 python scripts/kernel_ab.py --cases 720p,nv12 --strides 0 --seconds 1.5 > gpurun_out/r02_kernel_ab3.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r02_kernel_ab3.log'):
    if l.startswith('{'):
        d=json.loads(l); print(d['case'], d['tune'], f"{d['frames_per_s']/1e6:.3f} M  frac_alg {d['frac_alg']:.3f}  {d['fused_ms']:.4f} ms")
PY
