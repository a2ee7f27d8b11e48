for i in 1; do
for v in new fly1 tag16; do
  if [ $v = new ]; then unset FP4_B200_LIB; else export FP4_B200_LIB=/root/repo/torch_bnb_fp4_b200/variants/libfp4_b200.$v.so; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/ab_$v$i.json 2> gpurun_out/ab_$v$i.err
  python -c "
import json;d=json.loads(open('gpurun_out/ab_$v$i.json').read().strip().splitlines()[-1]);print('$v',$i,d['ms_per_step'],d['value'],d['nccl_allreduce_variant']['max_rel_diff_peer_vs_nccl'])"
done; done
