timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
bash tools/r02_gpu_profile.sh
