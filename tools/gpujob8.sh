set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tools/brillig_fast_path_bench.py 10000 1024 > gpurun_out/r2_brillig_fast_path.txt 2> gpurun_out/r2_brillig_fast_path.log
ACVMB_OPTS=scaled_columns=1 timeout 600 python tests/e2e_breakdown.py 2184 > gpurun_out/r2_e2e_breakdown_scaled1.txt 2>&1
ACVMB_OPTS=scaled_columns=0 timeout 600 python tests/e2e_breakdown.py 2184 > gpurun_out/r2_e2e_breakdown_scaled0.txt 2>&1
python tools/d2h_ceiling.py --gib 4 --reps 6 > gpurun_out/r2_d2h_ceiling_1gpu.txt 2>&1
