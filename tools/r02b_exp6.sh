set -x
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6
RLMPC_STEPS=120 timeout 900 python -m mpc4rl_b200.examples.cartpole_mpc_actor_critic 2>&1 | cut -c1-200 > gpurun_out/closed_loop_1gpu_120.log; awk 'NR%10==0' gpurun_out/closed_loop_1gpu_120.log | cut -c1-150
SKIPTESTS=1 WORKLOADS="cartpole cartpole_replay" bash tools/r02b_quick.sh r02f 2>&1 | grep "^r02f"
