export DVG_LIB_NOREBUILD=1
for t in v1 v2 v1 v2; do
DVG_LIB_TAG=$t timeout 200 python scripts/step_time.py --tag $t 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['tag'], [round(s['us_per_step_best'],2) for s in d['steps']])"
done
DVG_LIB_TAG=v1 DVG_STEP_CHAIN=0 timeout 200 python scripts/step_time.py --tag v1nochain 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['tag'], [round(s['us_per_step_best'],2) for s in d['steps']])"
DVG_LIB_TAG=v2 DVG_STEP_CHAIN=0 timeout 200 python scripts/step_time.py --tag v2nochain 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['tag'], [round(s['us_per_step_best'],2) for s in d['steps']])"
