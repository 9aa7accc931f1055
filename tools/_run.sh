set -x
mkdir -p gpurun_out
timeout 300 python tools/stage_times.py --workload c4 --iters 5 --out gpurun_out/r02g_stages_c4.txt > /dev/null 2> gpurun_out/r02g.err
timeout 300 python tools/stage_times.py --workload chig --out gpurun_out/r02g_stages_chig.txt > /dev/null 2>> gpurun_out/r02g.err
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02g_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02g_pytest_gpu.log
grep -E "passed|failed|FAILED|Error" gpurun_out/r02g_pytest_gpu.log | head
grep -E "edge_|sum|graph" gpurun_out/r02g_stages_c4.txt
tail -2 gpurun_out/r02g_stages_chig.txt
