set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_pytest_gpu_5.log
run() { name=$1; shift; ACVMB_OPTS=$1 timeout 900 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --secondary none $2 $3 > gpurun_out/r2_bench_$name.json 2> gpurun_out/r2_bench_$name.log; }
run fast
run fast_noir "" --coeffs noir-like
run fast_ring ring_bytes=24576
timeout 900 python bench.py --steps 1 --warmup 1 --no-e2e --secondary quick > gpurun_out/r2_bench_secondary_quick.json 2> gpurun_out/r2_bench_secondary_quick.log
