timeout 300 python -m pytest tests/test_gpu_hf.py -q -m gpu 2>&1 | tail -4
bash tools/run_tp.sh 2
