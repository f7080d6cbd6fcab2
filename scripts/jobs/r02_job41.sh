for t in x5 x6; do
DVG_LIB_TAG=$t timeout 200 python scripts/step_time.py --tag $t 2>&1 | tail -1
done
