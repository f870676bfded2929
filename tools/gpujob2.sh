set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_pytest_gpu_2.log
for sc in 1 0; do
  ACVMB_SCALED=$sc timeout 600 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_scaled$sc.json 2> gpurun_out/r2_bench_scaled$sc.log
done
ACVMB_SCALED=1 timeout 600 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --coeffs noir-like > gpurun_out/r2_bench_scaled1_noir.json 2> gpurun_out/r2_bench_scaled1_noir.log
