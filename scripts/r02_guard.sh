#!/bin/bash
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_gpu_edge.py -x -q -m gpu -k red_zone > gpurun_out/r02_guard_selftest.log 2>&1; echo "selftest rc=$?"; tail -3 gpurun_out/r02_guard_selftest.log
ESD_GUARD=1 timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r02_guard_suite.log 2>&1; echo "guarded suite rc=$?"; tail -4 gpurun_out/r02_guard_suite.log; grep -c "esd guard" gpurun_out/r02_guard_suite.log
ESD_GUARD=1 timeout 300 python scripts/sanitize_case.py 2>&1 | tail -3
