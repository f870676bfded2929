set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ACVMB_OPTS=slot_interleave=1 timeout 300 python tests/profile_target_hash.py 4096 > gpurun_out/r2_hash_il1.log 2>&1
ACVMB_OPTS=slot_interleave=0 timeout 300 python tests/profile_target_hash.py 4096 > gpurun_out/r2_hash_il0.log 2>&1
timeout 600 python -m pytest tests/test_gpu_blackbox.py tests/test_gpu_arith.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_pytest_gpu_18.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vm_kernel -c 1 -o gpurun_out/r2_hash_core_v7 -f python tests/profile_target_hash.py 256 > gpurun_out/r2_hash_ncu_v7.log 2>&1
