cd torch_bnb_fp4_b200
cp csrc/gemv_stream.cu /tmp/new_stream.cu
for v in new old new old; do
  if [ $v = old ]; then cp ../tools/_old_stream.cu.txt csrc/gemv_stream.cu; else cp /tmp/new_stream.cu csrc/gemv_stream.cu; fi
  touch csrc/gemv_stream.cu
  python build.py > /dev/null 2>&1
  echo "=== $v"
  python ../tools/microbench.py --batch 1 --no-dequant --shapes 4096x4096 14336x4096 4096x14336
done
cp /tmp/new_stream.cu csrc/gemv_stream.cu
