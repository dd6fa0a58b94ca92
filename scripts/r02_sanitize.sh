#!/bin/bash
cd "$(dirname "$0")/.."
timeout 300 python scripts/sanitize_case.py > gpurun_out/r02_sanitize_plain.log 2>&1; echo "plain rc=$?"; tail -4 gpurun_out/r02_sanitize_plain.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_case.py > gpurun_out/r02_sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -8 gpurun_out/r02_sanitize_memcheck.log
