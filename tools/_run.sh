set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02b_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest.log
timeout 300 python tools/stage_times.py --workload chig --opts fused=1 --out gpurun_out/r02b_stages_chig_fused.txt > /dev/null 2> gpurun_out/r02b_stages_fused.err
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/r02b_bench_chig.json 2> gpurun_out/r02b_bench_chig.err
tail -5 gpurun_out/r02b_pytest.log
tail -12 gpurun_out/r02b_stages_chig_fused.txt
tail -3 gpurun_out/r02b_bench_chig.err
