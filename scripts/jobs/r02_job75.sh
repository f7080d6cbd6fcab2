export DVG_LIB_NOREBUILD=1
fmt='import sys,json; d=json.loads(sys.stdin.read()); print(d["tag"], [round(s["us_per_step_best"],2) for s in d["steps"]])'
for t in base x3 base x3; do
DVG_LIB_TAG=$t timeout 200 python scripts/step_time.py --tag $t 2>&1 | tail -1 | python -c "$fmt"
done
