for v in 1 2; do
DVG_STEP_NOTRIG=$v timeout 200 python scripts/step_time.py --tag notrig$v 2>&1 | tail -1
done
