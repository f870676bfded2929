set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1500 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.log ) 2> gpurun_out/r2_bench_default.time
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.log ) 2> gpurun_out/r2_bench_reference.time
timeout 2400 python -m pytest tests/test_gpu_full_size.py -m gpu -x -q --durations=12 -k "config3 or config2 or config4 or other_modes" 2>&1 | tail -25 > gpurun_out/r2_pytest_full_size.log
timeout 300 python tests/profile_target_hash.py 1024 > gpurun_out/r2_hash_time2.log 2>&1
