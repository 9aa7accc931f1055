#!/bin/bash
# One record run on a GPU box: bench lines of every workload, ncu launch lists and --set full captures of the
# dominant kernels.  Usage (from the repo root):  gpurun --timeout 1500 -- 'bash tools/record_run.sh r01h'
# Outputs land in gpurun_out/<tag>_*; summarise them into profiles/ with tools/ncu_summary.py afterwards.
tag=${1:-rec}
out=gpurun_out
mkdir -p $out
python bench.py > $out/${tag}_bench_chig.json 2> $out/${tag}_bench_chig.err
for w in trpcage ww abd c4 c5; do
  timeout 300 python bench.py --workload $w --steps 30 --warmup 3 --skip-cpu-baseline > $out/${tag}_bench_$w.json 2> $out/${tag}_bench_$w.err
done
for w in chig c4; do
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_$w.csv \
    python bench.py --workload $w --steps 2 --warmup 1 --skip-cpu-baseline > $out/ncu_list_$w.log 2>&1
  for k in edge_fwd_tc edge_bwd_tc; do
    timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 14 -c 2 -f -o $out/${tag}_${k}_$w \
      python bench.py --workload $w --steps 2 --warmup 1 --skip-cpu-baseline > $out/ncu_full_${k}_$w.log 2>&1
  done
done
timeout 120 python tools/tc_timeline.py --workload chig > $out/${tag}_timeline_chig.txt 2>&1
timeout 120 python tools/tc_crossover.py --stages 19 > $out/${tag}_stages_19frag.txt 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > $out/${tag}_gpu.txt
ls -la $out | tail -40
