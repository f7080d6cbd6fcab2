timeout 600 python -m pytest tests/test_gpu_gp.py tests/test_gpu_rollout.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python scripts/gp_sweep.py --M 96,128,256 2>&1 | grep '^{'
