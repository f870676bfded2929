set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { name=$1; shift; ACVMB_OPTS=$1 timeout 900 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --secondary none $2 $3 $4 $5 > gpurun_out/r2_bench_$name.json 2> gpurun_out/r2_bench_$name.log; }
run sortk_noir "" --coeffs noir-like
run sortk_dense ""
run sortk_noir_global "" --coeffs noir-like --mode global
timeout 600 python -m pytest tests/test_gpu_arith.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2_pytest_arith_sortk.log
