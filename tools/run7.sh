timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/r02_pytest7.log 2>&1; tail -6 gpurun_out/r02_pytest7.log
timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02_bench_reference_ext.json 2> gpurun_out/r02_bench_ref.err; tail -2 gpurun_out/r02_bench_ref.err
timeout 1200 python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; tail -2 gpurun_out/r02_bench_1gpu.err
python - <<'PY'
import json
r=json.loads(open('gpurun_out/r02_bench_reference_ext.json').read().strip().splitlines()[-1])
d=json.loads(open('gpurun_out/r02_bench_1gpu.json').read().strip().splitlines()[-1])
print('ref', r['value'], r['tok_per_s'], r['e2e']['value'], r.get('clocks'))
print('ours', d['value'], d['tok_per_s'], d['e2e']['value'], d['config']['launches_per_step'], d['roofline']['frac'], d['clocks'])
print('ungrouped', d['ungrouped_launches']['value'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
print('sanity batch16 bf16', d['sanity_mlp']['bfloat16']['batch16'])
PY
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02_ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:gemv_stream -s 896 -c 256 --csv --log-file gpurun_out/r02_launches_raw.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02_ncu_list.log 2>&1
python tools/ncu_launch_list.py gpurun_out/r02_launches_raw.csv gpurun_out/r02_bench_launches.csv 128
