set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -22 > gpurun_out/r2_pytest_gpu_final.log
timeout 1500 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_arith.py tests/test_gpu_blackbox.py tests/test_reference_solver_cases.py tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_sanitizer_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_arith.py tests/test_gpu_blackbox.py -m gpu -x -q -k "synthetic_1k or mixed_failures or pedersen_chain or mixed_circuit or ring_of_recent or scaled_and_canonical" 2>&1 | tail -8 > gpurun_out/r2_sanitizer_racecheck.log
python __graft_entry__.py --smoke > gpurun_out/r2_smoke.log 2>&1
