set -x
cd torch_bnb_fp4_b200
touch csrc/gemv_stream.cu
FP4_B200_NVCC_EXTRA="-DFP4_STREAM_WARPS=16 -DFP4_STREAM_MINB=2" python build.py > /dev/null 2>&1
echo "=== WARPS=16 MINB=2 smem 112"
FP4_B200_GEMV_SMEM_KB=112 python ../tools/microbench.py --batch 1 --no-dequant
echo "=== WARPS=16 MINB=2 smem 226"
python ../tools/microbench.py --batch 1 --no-dequant
touch csrc/gemv_stream.cu
FP4_B200_NVCC_EXTRA="-DFP4_STREAM_WARPS=16 -DFP4_STREAM_MINB=1" python build.py > /dev/null 2>&1
echo "=== WARPS=16 MINB=1 smem 226"
python ../tools/microbench.py --batch 1 --no-dequant
