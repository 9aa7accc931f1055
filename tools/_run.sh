set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02f_gpus.txt
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -x -q > gpurun_out/r02f_pytest_multigpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02f_pytest_multigpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r02f_bench_chig_n2.json 2> gpurun_out/r02f_bench_chig_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 5 --nccl --no-c4 > gpurun_out/r02f_bench_chig_n2_nccl.json 2> gpurun_out/r02f_bench_chig_n2_nccl.err
tail -20 gpurun_out/r02f_pytest_multigpu.log
tail -3 gpurun_out/r02f_bench_chig_n2.err
head -c 1500 gpurun_out/r02f_bench_chig_n2.json
