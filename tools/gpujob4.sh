set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
# full-size gate kernel under ncu (one launch), ring off and on
ACVMB_OPTS=ring_bytes=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:vm_kernel -s 1 -c 1 -o gpurun_out/r2_vm_scaled_full python tests/profile_target.py 1048576 4736 2 > gpurun_out/r2_vm_scaled_full.log 2>&1
ACVMB_OPTS=ring_bytes=24576 timeout 900 ncu --set full --clock-control none -k regex:vm_kernel -s 1 -c 1 -o gpurun_out/r2_vm_scaled_ring python tests/profile_target.py 1048576 4736 2 > gpurun_out/r2_vm_scaled_ring.log 2>&1
run() { name=$1; shift; ACVMB_OPTS=$1 timeout 600 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --secondary none $2 $3 > gpurun_out/r2_bench_$name.json 2> gpurun_out/r2_bench_$name.log; }
run ring0 ring_bytes=0
run ring0_T16 ring_bytes=0,T=16
run ring0_chunk4 ring_bytes=0,chunk_steps=4
# multi-device C ABI smoke (1 GPU here: n = 1 path) + secondary quick
timeout 600 python bench.py --steps 1 --warmup 1 --no-e2e --secondary quick > gpurun_out/r2_bench_secondary_quick.json 2> gpurun_out/r2_bench_secondary_quick.log
