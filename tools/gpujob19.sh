set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vm_kernel -c 1 -o gpurun_out/r2_pedersen_chain_v7 -f python tests/profile_target_pedersen.py 32 > gpurun_out/r2_pedersen_ncu_v7.log 2>&1
timeout 300 python tests/profile_target_pedersen.py 256 > gpurun_out/r2_pedersen_256.log 2>&1
