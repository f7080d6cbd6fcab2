timeout 600 python -m pytest tests/test_gpu_gp.py -x -q -m gpu 2>&1 | tail -12
timeout 300 python scripts/gp_sweep.py --M 256,1024,4096 --samples 256,4096 2>&1 | tail -8
DVG_GP_TC=0 timeout 300 python scripts/gp_sweep.py --M 1024 --samples 4096 2>&1 | tail -2
