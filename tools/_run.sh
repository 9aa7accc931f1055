set -x
out=gpurun_out
mkdir -p $out
nvidia-smi -L > $out/r02w_gpus.txt
timeout 600 python tools/stage_times.py --workload chig 2>&1 | grep -E "embed_node|graph replay"
timeout 600 python -m pytest tests/test_stages_gpu.py -m gpu -q -x 2>&1 | tail -2
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -x -q > $out/r02w_pytest_multigpu_2gpu.log 2>&1
tail -3 $out/r02w_pytest_multigpu_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 > $out/r02w_bench_chig_n2.json 2> $out/r02w_bench_chig_n2.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02w_bench_chig_n2.json") if l.startswith("{")][0])
print("n2 value", d["value"], "e2e", d["e2e"]["value"], "md", (d.get("md_device") or {}).get("value"), "comm", d.get("comm"), "scale_c4", d.get("scale_c4"), "parity", (d.get("parity") or {}).get("parity_ok"))
PY
