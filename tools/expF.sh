{
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5
for pf in -1 0 1 2 8; do
  FP4_B200_GEMV_PF_UNITS=$pf python tools/stack_time.py --steps 20
done
} > gpurun_out/expF.log 2>&1
grep -E "passed|failed|error|ms/token" gpurun_out/expF.log
