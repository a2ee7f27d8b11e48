timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/r02_pytest6.log 2>&1; tail -12 gpurun_out/r02_pytest6.log
timeout 900 python bench.py --workload llama70b --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02_bench_llama70b_nested_1gpu.json 2> gpurun_out/r02_bench_llama70b_1gpu.err; tail -2 gpurun_out/r02_bench_llama70b_1gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_llama70b_nested_1gpu.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','tok_per_s')}, d['config']['launches_per_step'], d['roofline']['frac'])
PY
for pre in 1 2 3 4; do for ring in 2 3 4; do FP4_B200_GEMV_PRE_STEPS=$pre FP4_B200_GEMV_RING=$ring python tools/stack_time.py --steps 15 2>&1 | grep grouped; done; done
