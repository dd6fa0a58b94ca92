#!/bin/bash
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest6.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest6.log
tail -6 gpurun_out/r02_pytest6.log
{
python scripts/kernel_ab.py --strides 0 --seconds 1.5
python scripts/kernel_ab.py --cases 720p --strides 0 --seconds 1.5 --tune pipeline_stages=3
python scripts/kernel_ab.py --cases 720p --strides 0 --seconds 1.5 --tune rows_per_group=8
python scripts/kernel_ab.py --cases 720p --strides 0 --seconds 1.5 --tune rows_per_group=8 --tune pipeline_stages=3
python scripts/kernel_ab.py --cases 720p --strides 0 --seconds 1.5 --tune rows_per_stage=2 --tune pipeline_stages=3
python scripts/kernel_ab.py --cases 720p --strides 0 --seconds 1.5 --tune rows_per_stage=2 --tune pipeline_stages=4
python scripts/kernel_ab.py --cases 720p,1080p --strides 0 --seconds 1.5 --tune rows_per_stage=2 --tune pipeline_stages=4 --tune rows_per_group=8
} > gpurun_out/r02_kernel_ab2.log 2> gpurun_out/r02_kernel_ab2.err
echo "ab rc=$?"; python - <<'PY'
import json
for l in open('gpurun_out/r02_kernel_ab2.log'):
    if l.startswith('{'):
        d=json.loads(l); print(d['case'], d['tune'], f"{d['frames_per_s']/1e6:.3f} M  frac_alg {d['frac_alg']:.3f} fetched {d['frac_fetched']:.3f}  {d['fused_ms']:.4f} ms")
PY
tail -3 gpurun_out/r02_kernel_ab2.err
