DVG_SMALL_TRACE=1 timeout 120 python scripts/small_trace.py 2>&1 | tail -19
timeout 600 python -m pytest tests/test_gpu_lstm.py tests/test_gpu_rollout.py -x -q -m gpu 2>&1 | tail -4
timeout 120 python scripts/step_time.py --workload smmnist_b16 --tag small2 2>/dev/null | tail -1
