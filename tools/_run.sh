set -x
mkdir -p gpurun_out
bash tools/record_run.sh r02 > gpurun_out/r02_record.log 2>&1
tail -5 gpurun_out/r02_record.log
du -sh gpurun_out
