timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02_pytest2.log 2>&1; tail -25 gpurun_out/r02_pytest2.log
python tools/stack_time.py --steps 20 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 3 --cpu-seconds 3 > gpurun_out/r02_bench2.json 2> gpurun_out/r02_bench2.err; tail -3 gpurun_out/r02_bench2.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench2.json'))
print({k:d[k] for k in ('value','ms_per_step','tok_per_s','gpu_launches')}, d['config']['launches_per_step'], d['e2e'], d['roofline']['frac'], d['clocks'])
print('ungrouped', d['ungrouped_launches'])
for k in ('c1_single_layer','sanity_mlp','gemm_sweep'): print(k, json.dumps(d.get(k))[:1500])
PY
