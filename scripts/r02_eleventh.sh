#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python scripts/decode_probe.py --backends native --sessions 8,12,16 --batch 256 --frames 256 --no-cpu > gpurun_out/r02_decode_probe6.log 2> gpurun_out/r02_decode_probe6.err
echo "probe6 rc=$?"; cat gpurun_out/r02_decode_probe6.log; tail -3 gpurun_out/r02_decode_probe6.err
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --e2e-modes gather --decode-sessions 16 --compressed-passes 6 > gpurun_out/r02_bench_s16.log 2>&1
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_s16.log') if l.startswith('{')][-1])
c=d["e2e_compressed"]; print("s16", {k:v for k,v in c.items() if k not in ("note","cpu_arm","decoder")})
PY
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --e2e-modes gather --decode-sessions 8 --compressed-passes 6 > gpurun_out/r02_bench_s8.log 2>&1
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_s8.log') if l.startswith('{')][-1])
c=d["e2e_compressed"]; print("s8", {k:v for k,v in c.items() if k not in ("note","cpu_arm","decoder")})
PY
