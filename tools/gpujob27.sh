set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for cfg in "T=8,S=16" "T=16,S=8" "T=32,S=4" "T=32,S=8" "T=16,S=16"; do
  ACVMB_OPTS=$cfg timeout 300 python tests/profile_target_hash.py 4096 > gpurun_out/r2_hash_shape_$cfg.log 2>&1
done
