set -x
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q --timeout 600) > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
run() { (timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu "$@") 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('RES', sys.argv[1:], round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), round(d['roofline']['avg_kernel_ms'],4), round(d['roofline']['kernel_share_of_step'],3), d['clocks']['sm_mhz'], d['clocks']['reasons'], round(d['e2e']['value']))" "$@"; }
run --tune rows_per_group=8
run --tune rows_per_group=16
run --tune rows_per_group=24
run --tune rows_per_group=36
run --tune rows_per_group=48
run --tune rows_per_group=16 --tune pipeline_stages=3
run --tune rows_per_group=16 --tune pipeline_stages=6
run --tune rows_per_group=16 --tune ctas_per_sm=2
run --tune rows_per_group=16 --tune split_mode=2
run --tune rows_per_group=16 --frames-per-step 512
run --tune rows_per_group=16 --frames-per-step 4096
