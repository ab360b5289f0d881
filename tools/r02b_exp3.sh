set -x
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
for p in 5 0 3 8; do echo "== passes $p"; bash tools/r02b_variants.sh cartpole --opt ipm_passes=$p 2>&1 | head -2; done
bash tools/r02b_variants.sh cartpole
bash tools/r02b_variants.sh cartpole_tiny_pert
bash tools/r02b_variants.sh evaporation
bash tools/r02b_variants.sh cartpole_bx
