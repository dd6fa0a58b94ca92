#!/bin/bash
cd "$(dirname "$0")/.."
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest10.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest10.log
tail -6 gpurun_out/r02_pytest10.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_c.log 2> gpurun_out/r02_bench_n1_c.err
echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_n1_c.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_n1_c.log') if l.startswith('{')][-1])
print("value", d["value"], "frac", d["roofline"]["frac"], "parity", d["parity"]["bit_exact"], "e2e", d["e2e"]["mode"], round(d["e2e"]["value"]))
c=d["e2e_compressed"]; print({k:v for k,v in c.items() if k not in ("note","cpu_arm")}); print(c.get("cpu_arm",{}).get("value"), c.get("ratio_vs_cpu_arm"))
print("cpu_baseline", d["cpu_baseline"]["value"], d["clocks"])
PY
