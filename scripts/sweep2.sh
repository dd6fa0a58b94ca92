mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q --timeout 600) > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
ESD_LIB=$PWD/eioku_b200/libesd_w4.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "clip_golden or random_sizes or batching" 2>&1 | tail -2
run() { (timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu "$@") 2>&1 | tail -1 | python -c "import sys,json,os; d=json.loads(sys.stdin.read()); print('RES', os.environ.get('ESD_LIB','default')[-12:], sys.argv[1:], round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), round(d['roofline']['avg_kernel_ms'],4), d['clocks']['sm_mhz'], d['clocks']['reasons'])" "$@"; }
run
run --tune pipeline_stages=3
export ESD_LIB=$PWD/eioku_b200/libesd_w4.so
run
run --tune pipeline_stages=3
run --tune pipeline_stages=2
run --tune rows_per_group=8 --tune pipeline_stages=3
export ESD_LIB=$PWD/eioku_b200/libesd_w16.so
run
run --tune pipeline_stages=6
unset ESD_LIB
timeout 600 python bench.py --gpus 1 --steps 300 --warmup 5 > gpurun_out/bench_n1.log 2>&1; tail -1 gpurun_out/bench_n1.log | cut -c1-600
