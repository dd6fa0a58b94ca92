mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q --timeout 600) > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
run() { (timeout 300 python bench.py --steps 400 --warmup 5 --no-cpu --e2e-frames 256 "$@") 2>&1 | tail -1 | python -c "import sys,json,os; d=json.loads(sys.stdin.read()); print('RES', sys.argv[1:], round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), round(d['roofline']['avg_kernel_ms'],4), d['clocks']['sm_mhz'], d['clocks']['reasons'], round(d['e2e']['value']))" "$@"; }
run --tune rows_per_stage=1
run --tune rows_per_stage=2
run --tune rows_per_stage=2 --tune pipeline_stages=2
run --tune rows_per_stage=2 --tune pipeline_stages=4
run --tune rows_per_stage=4 --tune pipeline_stages=2
run --tune rows_per_stage=4 --tune pipeline_stages=3
run --tune rows_per_stage=2 --tune rows_per_group=8
run --tune rows_per_stage=2 --tune rows_per_group=24
