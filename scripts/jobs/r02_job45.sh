for x in 1 3 5 6; do
DVG_LIB_TAG=x$x DVG_STEP_X=$x timeout 200 python scripts/step_time.py --tag x$x 2>&1 | tail -1
done
