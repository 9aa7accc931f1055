set -x
mkdir -p gpurun_out
timeout 600 python tools/stage_times.py --workload chig 2>&1 | grep -E "head|graph replay"
timeout 600 python tools/stage_times.py --workload chig --max-frags 1 2>&1 | grep -E "head|graph replay"
timeout 900 python -m pytest tests/test_stages_gpu.py tests/test_engine_gpu.py -m gpu -q -x 2>&1 | tail -3
