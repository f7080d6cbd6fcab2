timeout 600 python -m pytest tests/test_gpu_gp.py -x -q -m gpu -k "native or large" 2>&1 | tail -3
timeout 500 python scripts/factor_time.py 2>&1 | tail -8 | tee gpurun_out/r02_factor_time.jsonl
