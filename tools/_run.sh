set -x
mkdir -p gpurun_out
python bench.py --steps 200 --warmup 10 > gpurun_out/r02q_bench_chig.json 2> gpurun_out/r02q_bench_chig.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02q_bench_chig.json") if l.startswith("{")][0])
print("value", d["value"], "ms", d["ms_per_step"], "warm", d.get("value_l2_warm"), "e2e", d["e2e"]["value"], "md", d["md_device"]["value"] if d.get("md_device") else None, "parity", d.get("parity"), "acc", d.get("accuracy"))
PY
for w in trpcage ww; do for o in node_tc=1 node_tc=0 node_tc=0,node_nb=4; do
  timeout 600 python tools/stage_times.py --workload $w --opts $o 2>&1 | grep -E "^workload|graph replay"
done; done
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
