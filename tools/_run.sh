set -x
out=gpurun_out
mkdir -p $out
timeout 600 python tools/stage_times.py --workload chig 2>&1 | grep -E "embed|graph replay"
cap() {
  rep=/tmp/s_$1
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$1 --launch-skip $2 -c 1 -f -o $rep \
    python bench.py --workload chig --steps 2 --warmup 1 --skip-cpu-baseline > $out/ncu_$1.log 2>&1
  python tools/ncu_summary.py full $rep.ncu-rep $out/r02t_$1_chig_full.txt > /dev/null 2>&1
  python tools/ncu_lines.py $rep.ncu-rep 16 > $out/r02t_$1_chig_lines.txt 2>&1
  rm -f $rep.ncu-rep
}
cap embed_node_small 3
cap embed_node_bwd 3
for k in embed_node_small embed_node_bwd; do
  echo "=== $k"; grep -E "duration|registers|warps_active|issue_active|stalled|inst_executed|dram|lts__t|l1tex" $out/r02t_${k}_chig_full.txt | head -16; head -16 $out/r02t_${k}_chig_lines.txt | cut -c1-170
done
