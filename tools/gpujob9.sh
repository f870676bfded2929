set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_pytest_multi_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 --gates 65536 --secondary quick > gpurun_out/r2_bench_2gpu_small.json 2> gpurun_out/r2_bench_2gpu_small.log
tail -5 gpurun_out/r2_bench_2gpu_small.log
timeout 600 python tools/brillig_fast_path_bench.py 10000 1024 > gpurun_out/r2_brillig_fast_path.txt 2> gpurun_out/r2_brillig_fast_path.log
