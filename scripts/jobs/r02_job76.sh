timeout 600 python -m pytest tests/test_gpu_gp.py -x -q -m gpu 2>&1 | tail -8
