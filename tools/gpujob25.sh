set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ACVMB_OPTS=sha_pad_table=1 timeout 300 python tests/profile_target_hash.py 4096 > gpurun_out/r2_hash_pad1.log 2>&1
ACVMB_OPTS=sha_pad_table=0 timeout 300 python tests/profile_target_hash.py 4096 > gpurun_out/r2_hash_pad0.log 2>&1
timeout 600 python -m pytest tests/test_gpu_blackbox.py -m gpu -x -q -k "hash" 2>&1 | tail -5 > gpurun_out/r2_pytest_gpu_25.log
