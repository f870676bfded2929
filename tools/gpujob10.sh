set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo_8gpu.txt 2>&1
timeout 300 python tools/d2h_ceiling.py --gib 2 --reps 6 > gpurun_out/r2_d2h_ceiling.txt 2>&1
timeout 300 python tools/d2h_ceiling.py --gib 2 --reps 6 --no-bind >> gpurun_out/r2_d2h_ceiling.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_pytest_multi_8gpu.log
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/r2_bench_8gpu.log ) 2> gpurun_out/r2_bench_8gpu.time
