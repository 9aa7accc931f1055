set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_pytest.log
timeout 300 python tools/stage_times.py --workload chig --opts node_tc=1 --out gpurun_out/r02d_stages_chig_nodetc.txt > /dev/null 2> gpurun_out/r02d_a.err
timeout 300 python tools/stage_times.py --workload c4 --opts node_tc=0 --iters 5 --out gpurun_out/r02d_stages_c4_simt.txt > /dev/null 2> gpurun_out/r02d_b.err
timeout 300 python tools/stage_times.py --workload c4 --opts node_tc=1 --iters 5 --out gpurun_out/r02d_stages_c4_nodetc.txt > /dev/null 2> gpurun_out/r02d_c.err
tail -15 gpurun_out/r02d_pytest.log
tail -4 gpurun_out/r02d_stages_chig_nodetc.txt gpurun_out/r02d_stages_c4_simt.txt gpurun_out/r02d_stages_c4_nodetc.txt
