set -x
mkdir -p gpurun_out
for w in chig c4; do
for o in krot=1 krot=0; do
  timeout 600 python tools/stage_times.py --workload $w --opts $o 2>&1 | grep -E "^workload|edge_fwd[03]|edge_bwd[03]|graph replay"
done; done > gpurun_out/r02p_krot.txt 2>&1
cat gpurun_out/r02p_krot.txt
timeout 900 python -m pytest tests/test_stages_gpu.py tests/test_engine_gpu.py -m gpu -q -x 2>&1 | tail -5
