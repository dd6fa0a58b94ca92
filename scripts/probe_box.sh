#!/bin/bash
# What the GPU box offers for decode (N1): video/jpeg driver libraries, engines, host cores.
{
echo "== ldconfig"; ldconfig -p | grep -i -E "nvcuvid|nvidia-encode|nvjpeg|libcuda\.so|nvidia-ml" 
echo "== find"; find / \( -name "libnvcuvid*" -o -name "libnvidia-encode*" -o -name "libnvjpeg*" \) -not -path "/proc/*" 2>/dev/null
echo "== caps env"; env | grep -i NVIDIA
echo "== smi"; nvidia-smi --query-gpu=name,driver_version,memory.total,pcie.link.gen.current,pcie.link.width.current --format=csv
nvidia-smi -q | grep -i -E -A4 "utilization|encoder|decoder|jpeg|ofa" | head -60
echo "== host"; nproc; free -g | head -2; lscpu | grep -E "Model name|Socket|NUMA|Thread|Core" 
python - <<'PY'
import ctypes
for n in ("libnvcuvid.so.1","libnvcuvid.so","libnvjpeg.so.12","libnvidia-encode.so.1"):
    try:
        ctypes.CDLL(n); print("dlopen ok", n)
    except OSError as e:
        print("dlopen FAIL", n, str(e)[:80])
PY
} > gpurun_out/probe_box.log 2>&1
exit 0
