# usage: run_tp.sh N [workload]   (gpurun --gpus N: the tensor-parallel bench line, summary printed; N=2 also runs tests/test_gpu_tp.py)
N=$1; WL=${2:-mistral7b}
if [ "$N" = "2" ]; then timeout 600 python -m pytest tests/test_gpu_tp.py -q -m gpu 2>&1 | tail -5; fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 30 --warmup 5 --workload $WL > gpurun_out/r02_bench_tp${N}_${WL}.json 2> gpurun_out/r02_bench_tp${N}_${WL}.err
tail -3 gpurun_out/r02_bench_tp${N}_${WL}.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_tp${N}_${WL}.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','tok_per_s','n_gpus','gpu_launches','scaling')}, d['config']['launches_per_step'])
print('e2e', d['e2e']); print('nccl', d.get('nccl_allreduce_variant')); print(d['tp_collective'][:80])
PY
