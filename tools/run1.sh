timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest1.log 2>&1; tail -15 gpurun_out/r02_pytest1.log
bash tools/ncu_one.sh r02_stream_4096 gemv_stream 40 4096x4096
