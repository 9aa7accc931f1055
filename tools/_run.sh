set -x
mkdir -p gpurun_out
timeout 600 python tools/stage_times.py --workload chig > gpurun_out/r02l_stages_chig.txt 2>&1
grep -E "node_fwd[036]|node_bwd[036]|graph replay|sum" gpurun_out/r02l_stages_chig.txt
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02l_pytest_gpu.log 2>&1
tail -8 gpurun_out/r02l_pytest_gpu.log
