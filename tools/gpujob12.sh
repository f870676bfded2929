set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -22 > gpurun_out/r2_pytest_gpu_final.log
