timeout 600 python -m pytest tests/test_gpu_gp.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python scripts/gp_sweep.py 2>&1 | grep '^{' > gpurun_out/r02_gp_sweep.jsonl; cat gpurun_out/r02_gp_sweep.jsonl
