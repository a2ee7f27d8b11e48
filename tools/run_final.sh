# final single-GPU validation of a round: GPU tests, smoke, both bench arms
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/r02_pytest_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; tail -2 gpurun_out/r02_bench_final.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_final.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['clocks'], d['cpu_baseline'])
PY
