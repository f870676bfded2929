set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_pytest_gpu_1.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:vm_kernel -c 1 -o gpurun_out/r2_hash_chain python tests/profile_target_hash.py 256 > gpurun_out/r2_hash_prof.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:vm_kernel -c 1 -o gpurun_out/r2_ped_chain python tests/profile_target_pedersen.py 32 > gpurun_out/r2_ped_prof.log 2>&1
timeout 300 python tests/profile_target_hash.py 1024 > gpurun_out/r2_hash_time.log 2>&1
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
lscpu > gpurun_out/lscpu.txt 2>&1
numactl -H > gpurun_out/numa.txt 2>&1
free -g >> gpurun_out/numa.txt
