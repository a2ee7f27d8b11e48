# one ncu --set full capture of eight consecutive launches (two decoder blocks) of the headline bench
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02_ncu_bench_plain.json 2> gpurun_out/r02_ncu_bench_plain.err || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemv_stream -s 896 -c 8 -f -o gpurun_out/r02_bench_stream_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02_ncu_bench.log 2>&1
tail -3 gpurun_out/r02_ncu_bench.log
