# ncu evidence for profiles/ (run under gpurun, 1 GPU). Plain run first, then the two ncu passes.
set -e
mkdir -p gpurun_out
CMD="python bench.py --steps 4 --warmup 3 --no-cpu --e2e-frames 256 --e2e-gather-threads 0"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fused_score -s 3 -c 2 -o gpurun_out/prof_fused -f $CMD > gpurun_out/ncu2.log 2>&1
echo "full rc=$?"
