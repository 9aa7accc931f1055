set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02i_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02i_pytest_gpu.log
grep -E "passed|failed|FAILED|Error" gpurun_out/r02i_pytest_gpu.log | head
for w in chig trpcage ww abd; do timeout 200 python tools/stage_times.py --workload $w --out gpurun_out/r02i_stages_$w.txt > /dev/null 2>> gpurun_out/r02i.err; done
timeout 300 python tools/stage_times.py --workload c4 --iters 5 --out gpurun_out/r02i_stages_c4.txt > /dev/null 2>> gpurun_out/r02i.err
for f in gpurun_out/r02i_stages_*.txt; do echo $f; head -1 $f | cut -c1-200; tail -n 2 $f; done
grep -E "embed|head" gpurun_out/r02i_stages_c4.txt
