#!/bin/bash
# One record run on a GPU box: bench lines of every workload, ncu launch lists and --set full captures of the
# dominant kernels.  Usage (from the repo root):  gpurun --timeout 2400 -- 'bash tools/record_run.sh r02'
# Outputs land in gpurun_out/<tag>_*; summarise them into profiles/ with tools/ncu_summary.py afterwards.
tag=${1:-rec}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > $out/${tag}_gpu.txt
timeout 600 python bench.py > $out/${tag}_bench_chig.json 2> $out/${tag}_bench_chig.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $out/${tag}_bench_chig_reference.json 2> $out/${tag}_bench_chig_reference.err
for w in trpcage ww abd c4 c5; do
  timeout 400 python bench.py --workload $w --steps 30 --warmup 3 --skip-cpu-baseline > $out/${tag}_bench_$w.json 2> $out/${tag}_bench_$w.err
done
timeout 120 python tools/stage_times.py --workload chig --out $out/${tag}_stages_19frag.txt > /dev/null 2>&1
timeout 120 python tools/stage_times.py --workload chig --max-frags 1 --out $out/${tag}_stages_1frag.txt > /dev/null 2>&1
timeout 120 python tools/stage_times.py --workload c4 --iters 5 --out $out/${tag}_stages_c4.txt > /dev/null 2>&1
timeout 120 python tools/tc_timeline.py --workload chig > $out/${tag}_timeline_chig.txt 2>&1
# --set full captures are summarised ON THE BOX (metrics + top source lines) and the reports deleted: gpurun brings back
# at most 64 MiB
capture() {   # capture <kernel regex> <workload> <launch-skip> <count>
  rep=/tmp/${tag}_$1_$2
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$1 --launch-skip $3 -c $4 -f -o $rep \
    python bench.py --workload $2 --steps 2 --warmup 1 --skip-cpu-baseline > $out/ncu_full_$1_$2.log 2>&1
  python tools/ncu_summary.py full $rep.ncu-rep $out/${tag}_$1_$2_full.txt > /dev/null 2>&1
  python tools/ncu_lines.py $rep.ncu-rep 25 > $out/${tag}_$1_$2_lines.txt 2>&1
  rm -f $rep.ncu-rep
}
for w in chig c4; do
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_$w.csv \
    python bench.py --workload $w --steps 2 --warmup 1 --skip-cpu-baseline > $out/ncu_list_$w.log 2>&1
  capture edge_fwd_tc $w 14 2
  capture edge_bwd_tc $w 14 2
done
capture node_tc_kernel c4 20 4
capture node_bwd2 chig 8 2
capture node_fwd2 chig 8 2
capture embed_node_small chig 3 1
capture embed_node_bwd chig 3 1
for t in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $t python tools/sanitize_run.py 3 4 node_tc=1 > $out/${tag}_${t}_nodetc.log 2>&1
done
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_run.py 3 4 caph=1 > $out/${tag}_memcheck_caph_md.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_run.py 3 4 fused=1 > $out/${tag}_memcheck_fused.log 2>&1
ls -la $out | tail -60
