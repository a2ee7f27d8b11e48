set -x
cd torch_bnb_fp4_b200
touch csrc/gemv_stream.cu
FP4_B200_NVCC_EXTRA="-DFP4_STREAM_WARPS=8 -DFP4_STREAM_MINB=2" python build.py > /dev/null 2>&1
echo "=== WARPS=8 x 2 CTAs/SM"
python ../tools/microbench.py --batch 1 --no-dequant
cd .. && python bench.py --steps 20 --no-cpu-baseline > gpurun_out/z5_bench.json 2> gpurun_out/z5_bench.err; python -c "
import json; b=json.load(open('gpurun_out/z5_bench.json')); print(b['value'], b['tok_per_s'], b['grouped_launches']['value'], b['grouped_launches']['tok_per_s'])"
