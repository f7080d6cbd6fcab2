timeout 600 python -m pytest tests/test_gpu_rollout.py -x -q -m gpu 2>&1 | tail -8
timeout 200 python scripts/step_time.py --tag new 2>&1 | tail -1
timeout 200 python scripts/step_time.py --workload bair_s32 --tag new 2>&1 | tail -1
DVG_STEP_CHAIN=0 timeout 200 python scripts/step_time.py --tag new_nochain 2>&1 | tail -1
