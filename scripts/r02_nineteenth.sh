#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r02_pytest19.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest19.log
for st in 1 8; do
ESD_GATHER_STREAMS=$st timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --no-compressed > gpurun_out/r02_gs$st.log 2> gpurun_out/r02_gs$st.err; echo "bench streams=$st rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02_gs$st.log') if l.startswith('{')][-1])
print("streams $st: value", round(d["value"]), "e2e", d["e2e"]["mode"], round(d["e2e"]["value"]), {k: round(v["value"]) for k,v in d["e2e"]["modes"].items()})
PY
done
ESD_GATHER_STREAMS=8 ESD_GATHER_PF8=768 timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --no-compressed > gpurun_out/r02_gs8b.log 2> gpurun_out/r02_gs8b.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02_gs8b.log') if l.startswith('{')][-1])
print("streams 8 pf 768: e2e", d["e2e"]["mode"], round(d["e2e"]["value"]), {k: round(v["value"]) for k,v in d["e2e"]["modes"].items()})
PY
