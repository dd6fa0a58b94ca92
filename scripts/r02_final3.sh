#!/bin/bash
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest_final3.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_final3.log
ESD_GUARD=1 timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02_guard_suite.log 2>&1; echo "guarded suite rc=$?"; tail -4 gpurun_out/r02_guard_suite.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_final3_ref.log 2> gpurun_out/r02_final3_ref.err; echo "ref rc=$?"; tail -c 400 gpurun_out/r02_final3_ref.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_final3_n1.log 2> gpurun_out/r02_final3_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_final3_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_final3_n1.log') if l.startswith('{')][-1])
print("value", d["value"], "frac", d["roofline"]["frac"], "parity", d["parity"]["bit_exact"], "e2e", d["e2e"]["mode"], round(d["e2e"]["value"]))
c=d["e2e_compressed"]; print({k:v for k,v in c.items() if k not in ("note","cpu_arm","decoder")}); print(c.get("cpu_arm",{}).get("value"), c.get("ratio_vs_cpu_arm"))
print("cpu_baseline", d["cpu_baseline"]["value"], d["clocks"])
PY
