set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1500 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.log ) 2> gpurun_out/r2_bench_final.time
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_final.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --secondary none --gates 65536 > gpurun_out/r2_launches_bench.log 2>&1
