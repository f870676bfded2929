set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ACVMB_OPTS=slack_scheduling=0 timeout 300 python tests/profile_target_pedersen.py 256 > gpurun_out/r2_pedersen_256_slack0.log 2>&1
ACVMB_OPTS=slack_scheduling=1 timeout 300 python tests/profile_target_pedersen.py 256 > gpurun_out/r2_pedersen_256_slack1.log 2>&1
ACVMB_OPTS=slack_scheduling=0,S=16 timeout 300 python tests/profile_target_pedersen.py 256 > gpurun_out/r2_pedersen_256_slack0_S16.log 2>&1
ACVMB_OPTS=slack_scheduling=1,S=16 timeout 300 python tests/profile_target_pedersen.py 256 > gpurun_out/r2_pedersen_256_slack1_S16.log 2>&1
