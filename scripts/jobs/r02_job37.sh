timeout 600 python -m pytest tests/test_gpu_rollout.py tests/test_gpu_lstm.py -x -q -m gpu 2>&1 | tail -8
for c in 1 0; do
  DVG_STEP_CHAIN=$c timeout 200 python scripts/step_time.py --tag chain$c 2>&1 | tail -1
  DVG_STEP_CHAIN=$c timeout 200 python scripts/step_time.py --workload bair_s32 --tag chain$c 2>&1 | tail -1
done
