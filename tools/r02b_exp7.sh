set -x
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
for c in 4 3; do echo "== coop blocks per SM $c"; RLMPC_COOP_BLOCKS_PER_SM=$c bash tools/r02b_variants.sh cartpole 2>&1 | grep "ms/step"; RLMPC_COOP_BLOCKS_PER_SM=$c bash tools/r02b_variants.sh cartpole_tiny_pert 2>&1 | grep "ms/step"; done
