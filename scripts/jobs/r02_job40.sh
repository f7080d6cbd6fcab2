DVG_STEP_CHAIN_VERBOSE=1 timeout 200 python scripts/step_time.py --steps 6 --reps 2 2>&1 | grep -v "^{" | sort | uniq -c | sort -rn | head -20
