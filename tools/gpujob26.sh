set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_blackbox.py -m gpu -x -q -k "hash or pedersen_chain or curve_ops_every_tile_shape_and_lowering" 2>&1 | tail -12 > gpurun_out/r2_sanitizer_memcheck_final.log
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_blackbox.py -m gpu -x -q -k "hash_calls_over_byte or hash_circuit_mixed" 2>&1 | tail -8 > gpurun_out/r2_sanitizer_racecheck_final.log
