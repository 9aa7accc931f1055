set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02k_pytest_gpu.log 2>&1
tail -25 gpurun_out/r02k_pytest_gpu.log
