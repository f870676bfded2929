set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_pytest_gpu_3.log
run() { name=$1; shift; ACVMB_OPTS=$1 timeout 600 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline $2 $3 > gpurun_out/r2_bench_$name.json 2> gpurun_out/r2_bench_$name.log; }
run ring24k ring_bytes=24576
run ring16k ring_bytes=16384
run ring32k_stage3 ring_bytes=32768,n_stage=3
run ring24k_noir ring_bytes=24576 --coeffs noir-like
run ring24k_global ring_bytes=24576 --mode global
run ring0_global ring_bytes=0 --mode global
