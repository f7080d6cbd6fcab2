DVG_LIB_TAG=x3 timeout 200 python scripts/step_time.py --tag x3 2>&1 | tail -1
