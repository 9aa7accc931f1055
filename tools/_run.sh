set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02e_pytest.log
for w in trpcage ww abd; do
  for o in 0 1; do
    timeout 300 python tools/stage_times.py --workload $w --opts node_tc=$o --out gpurun_out/r02e_stages_${w}_nodetc$o.txt > /dev/null 2>> gpurun_out/r02e.err
  done
done
timeout 300 python tools/stage_times.py --workload c4 --opts node_tc=1 --iters 5 --out gpurun_out/r02e_stages_c4_nodetc.txt > /dev/null 2>> gpurun_out/r02e.err
timeout 300 python tools/stage_times.py --workload chig --out gpurun_out/r02e_stages_chig.txt > /dev/null 2>> gpurun_out/r02e.err
tail -25 gpurun_out/r02e_pytest.log
for f in gpurun_out/r02e_stages_*.txt; do echo $f; tail -n 2 $f; done
