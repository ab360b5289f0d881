timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3
SKIPTESTS=1 WORKLOADS="cartpole cartpole_tiny_pert cartpole_bx" bash tools/r02b_quick.sh r02m 2>&1 | grep "^r02m" | cut -c1-330
