# Tuning sweep of the fused kernel (run under gpurun on one B200); results are recorded in profiles/r01_sweep.md.
run() { (timeout 300 python bench.py --steps 400 --warmup 5 --no-cpu --e2e-frames 128 "$@") 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('RES', sys.argv[1:], round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), round(d['roofline']['avg_kernel_ms'],4), d['clocks']['sm_mhz'], d['clocks']['reasons'])" "$@"; }
run
for rs in 1 2 4; do for st in 2 3 4; do run --tune rows_per_stage=$rs --tune pipeline_stages=$st; done; done
for r in 8 16 24; do run --tune rows_per_group=$r; done
run --tune split_mode=2
run --tune ctas_per_sm=2
