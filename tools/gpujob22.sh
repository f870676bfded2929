set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tests/profile_target_pedersen.py 256 > gpurun_out/r2_pedersen_256_slack1.log 2>&1
timeout 900 python -m pytest tests/test_gpu_blackbox.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_pytest_gpu_22.log
timeout 900 python -m pytest tests/test_gpu_full_size.py -m gpu -x -q -k "config2 or config4" 2>&1 | tail -5 > gpurun_out/r2_pytest_full_22.log
