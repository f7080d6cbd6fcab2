DVG_LIB_TAG=x5 DVG_STEP_NOTRIG=1 timeout 200 python scripts/step_time.py --tag x5notrig1 2>&1 | tail -1
