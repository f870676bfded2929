set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_blackbox.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_pytest_gpu_17.log
ACVMB_OPTS=packed_hashes=1 timeout 300 python tests/profile_target_hash.py 4096 > gpurun_out/r2_hash_packed1.log 2>&1
ACVMB_OPTS=packed_hashes=0 timeout 300 python tests/profile_target_hash.py 4096 > gpurun_out/r2_hash_packed0.log 2>&1
timeout 600 python -m pytest tests/test_gpu_full_size.py -m gpu -x -q -k "config3 or config4" 2>&1 | tail -5 > gpurun_out/r2_pytest_full_17.log
