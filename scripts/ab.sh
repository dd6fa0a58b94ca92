run() { d=$1; shift; (cd $d && timeout 300 python bench.py --steps 400 --warmup 5 --no-cpu --e2e-frames 256 "$@") 2>&1 | tail -1 | python -c "import sys,json,os; d=json.loads(sys.stdin.read()); print('RES', sys.argv[1:], round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'], d['clocks']['reasons'], d['clocks'].get('power_w_max'))" $d "$@"; }
nvidia-smi --query-gpu=name,power.limit,clocks.max.sm,temperature.gpu --format=csv
for i in 1 2 3; do
run _old
run .
run . --tune rows_per_stage=1
run . --tune rows_per_stage=2 --tune pipeline_stages=2
run . --tune rows_per_stage=4 --tune pipeline_stages=2
done
