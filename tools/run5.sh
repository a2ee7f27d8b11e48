timeout 900 python bench.py --workload llama70b --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02_bench_llama70b_nested_1gpu.json 2> gpurun_out/r02_bench_llama70b_1gpu.err; tail -2 gpurun_out/r02_bench_llama70b_1gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_llama70b_nested_1gpu.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','tok_per_s')}, d['config']['launches_per_step'], d['config']['workload'][:90], d['roofline']['frac'])
print('ungrouped', d['ungrouped_launches'])
PY
timeout 900 python bench.py --workload llama70b --no-nested --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02_bench_llama70b_fp32absmax_1gpu.json 2>> gpurun_out/r02_bench_llama70b_1gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_llama70b_fp32absmax_1gpu.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','tok_per_s')}, d['config']['launches_per_step'], d['roofline']['frac'])
PY
python tools/sanitize_case.py > gpurun_out/r02_sanitize_plain.log 2>&1 && compute-sanitizer --tool memcheck --log-file gpurun_out/r02_sanitizer_memcheck.log python tools/sanitize_case.py > gpurun_out/r02_sanitize_memcheck_stdout.log 2>&1
tail -3 gpurun_out/r02_sanitize_plain.log; tail -5 gpurun_out/r02_sanitizer_memcheck.log
