# usage: ncu_one.sh <tag> <kernel-regex> <skip> <shape...>   (development: one ncu --set full capture of the microbenchmark)
tag=$1; rx=$2; skip=$3; shift 3
python tools/microbench.py --no-dequant --shapes "$@" > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 2 -f -o gpurun_out/$tag python tools/microbench.py --no-dequant --shapes "$@" > gpurun_out/${tag}_ncu.log 2>&1
cat gpurun_out/${tag}_plain.log
