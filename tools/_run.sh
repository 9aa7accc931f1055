set -x
out=gpurun_out
tag=r02x
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > $out/${tag}_gpu.txt
timeout 600 python bench.py > $out/${tag}_bench_chig.json 2> $out/${tag}_bench_chig.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $out/${tag}_bench_chig_reference.json 2> $out/${tag}_bench_chig_reference.err
timeout 120 python tools/stage_times.py --workload chig --out $out/${tag}_stages_19frag.txt > /dev/null 2>&1
timeout 120 python tools/stage_times.py --workload chig --max-frags 1 --out $out/${tag}_stages_1frag.txt > /dev/null 2>&1
timeout 120 python tools/stage_times.py --workload c4 --iters 5 --out $out/${tag}_stages_c4.txt > /dev/null 2>&1
timeout 120 python tools/tc_timeline.py --workload chig > $out/${tag}_timeline_chig.txt 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_chig.csv \
    python bench.py --workload chig --steps 2 --warmup 1 --skip-cpu-baseline > $out/ncu_list_chig.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; tail -2 $out/${tag}_smoke.log
tail -3 $out/${tag}_stages_19frag.txt; tail -2 $out/${tag}_stages_1frag.txt
