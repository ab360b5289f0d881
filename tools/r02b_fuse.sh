echo "== baseline (separate k_lin)"; bash tools/r02b_variants.sh cartpole_tiny_pert 2>&1 | sed -n 2p
echo "== fuse_lin=1"; bash tools/r02b_variants.sh cartpole_tiny_pert --opt fuse_lin=1
echo "== headline fuse_lin=1"; bash tools/r02b_variants.sh cartpole --opt fuse_lin=1 | sed -n 1,2p
timeout 600 python -m pytest tests/test_gpu_cartpole.py -q -m gpu 2>&1 | tail -2
