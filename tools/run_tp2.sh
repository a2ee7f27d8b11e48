# 2-GPU validation (gpurun --gpus 2): the peer-exchange check with its log kept, then the TP tests and the N=2 bench
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/tp_fused_check.py > gpurun_out/r02_tp_fused_check.log 2>&1
grep -v Warning gpurun_out/r02_tp_fused_check.log | grep "rank 0\|rank 1" | grep -v "done\|ready\|built" | tail -12
bash tools/run_tp.sh 2
