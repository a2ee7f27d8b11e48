cd torch_bnb_fp4_b200
touch csrc/gemv_stream.cu
FP4_B200_NVCC_EXTRA="-DFP4_STREAM_TIMELINE" python build.py > /dev/null 2>&1
cd ..
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 tools/tp_timeline.py 2>&1 | grep -E "launch|min" 
