DVG_STEP_NOTRIG=1 timeout 200 python scripts/step_time.py --tag notrig1 2>&1 | tail -12
