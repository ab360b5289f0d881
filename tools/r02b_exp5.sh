set -x
RLMPC_STEPS=200 timeout 900 python -m mpc4rl_b200.examples.cartpole_mpc_actor_critic 2>&1 | cut -c1-260 > gpurun_out/closed_loop_1gpu_200.log; awk 'NR%10==0' gpurun_out/closed_loop_1gpu_200.log | cut -c1-200
SKIPTESTS=1 WORKLOADS="evaporation cartpole_bx" bash tools/r02b_quick.sh r02e 2>&1 | grep "^r02e"
