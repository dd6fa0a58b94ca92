#!/bin/bash
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest18.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest18.log
timeout 200 python scripts/decide_probe.py > gpurun_out/r02_decide_probe2.log 2>&1; cat gpurun_out/r02_decide_probe2.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu > gpurun_out/r02_n1_lanes.log 2> gpurun_out/r02_n1_lanes.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_n1_lanes.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_n1_lanes.log') if l.startswith('{')][-1])
print("value", d["value"], "frac", d["roofline"]["frac"], "parity", d["parity"]["bit_exact"], "e2e", d["e2e"]["mode"], round(d["e2e"]["value"]))
c=d["e2e_compressed"]; print({k:v for k,v in c.items() if k not in ("note","cpu_arm","decoder")})
PY
