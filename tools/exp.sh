set -x
cd torch_bnb_fp4_b200
touch csrc/gemv_i8.cu
FP4_B200_NVCC_EXTRA="-DFP4_I8_NOCOMPUTE -DFP4_I8_TIMELINE" python build.py > /dev/null 2>&1
export FP4_B200_GEMV_CTAS_PER_SM=1
KW=16 python ../tools/i8_timeline.py 14336 4096 5 | tail -24
python ../tools/i8_timeline.py 4096 4096 5 | tail -16
touch csrc/gemv_i8.cu
FP4_B200_NVCC_EXTRA="-DFP4_I8_TIMELINE" python build.py > /dev/null 2>&1
python ../tools/i8_timeline.py 14336 4096 5 | tail -16
