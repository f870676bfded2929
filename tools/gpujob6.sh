set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { name=$1; shift; ACVMB_OPTS=$1 timeout 900 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --secondary none $2 $3 > gpurun_out/r2_bench_$name.json 2> gpurun_out/r2_bench_$name.log; }
run rev
run rev_chunk4 chunk_steps=4
run rev_noir "" --coeffs noir-like
timeout 2400 python -m pytest tests/test_gpu_full_size.py -m gpu -x -q --durations=10 2>&1 | tail -25 > gpurun_out/r2_pytest_full_size.log
