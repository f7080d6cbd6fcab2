timeout 600 python -m pytest tests/test_gpu_lstm.py tests/test_gpu_rollout.py -x -q -m gpu 2>&1 | tail -8
timeout 120 python scripts/step_time.py --workload smmnist_b16 --tag small_trig 2>/dev/null | tail -1
