timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/r02_pytest4.log 2>&1; tail -12 gpurun_out/r02_pytest4.log
python tools/stack_time.py --steps 20 2>&1 | tail -2
python tools/microbench.py --no-dequant --shapes 4096x4096 14336x4096 1024x4096 2>&1 | tail -3
