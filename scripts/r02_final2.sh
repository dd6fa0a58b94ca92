#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_final2_ref.log 2> gpurun_out/r02_final2_ref.err; echo "ref rc=$?"; tail -c 500 gpurun_out/r02_final2_ref.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_final2_n1.log 2> gpurun_out/r02_final2_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_final2_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_final2_n1.log') if l.startswith('{')][-1])
print("value", d["value"], "frac", d["roofline"]["frac"], "parity", d["parity"]["bit_exact"], "e2e", d["e2e"]["mode"], round(d["e2e"]["value"]))
c=d["e2e_compressed"]; print({k:v for k,v in c.items() if k not in ("note","cpu_arm","decoder")}); print(c.get("cpu_arm",{}).get("value"), c.get("ratio_vs_cpu_arm"))
print("cpu_baseline", d["cpu_baseline"]["value"], d["clocks"])
PY
bash scripts/gpu_profile.sh r02
