# ncu evidence for profiles/ (run under gpurun, 1 GPU). Plain run first, then the ncu pass, as B200_PROFILING.md asks.
# usage: bash scripts/gpu_profile.sh [tag]    -> gpurun_out/<tag>_launches.csv (+ scripts/ncu_summary.py <tag> for the .md)
TAG=${1:-r02}
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --min-segments 2 --min-seconds 0 --no-cpu --no-compressed --e2e-frames 256 --e2e-modes dma_rows --e2e-steps 2"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
echo "launch list rc=$?"
cp gpurun_out/launches.csv gpurun_out/${TAG}_launches.csv 2>/dev/null
