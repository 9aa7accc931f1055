set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_pytest.log
timeout 120 python tools/tc_selftest.py > gpurun_out/r02c_tc_selftest.txt 2>&1
tail -15 gpurun_out/r02c_pytest.log
tail -3 gpurun_out/r02c_tc_selftest.txt
