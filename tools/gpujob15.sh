set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_blackbox.py tests/test_gpu_arith.py tests/test_reference_solver_cases.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_pytest_gpu_15.log
