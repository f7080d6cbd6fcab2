timeout 900 python -m pytest tests/test_gpu_lstm.py tests/test_gpu_rollout.py tests/test_gpu_gp.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python scripts/hidden_sweep.py 2>/dev/null | tee gpurun_out/r02_hidden_sweep.jsonl
