export DVG_LIB_NOREBUILD=1
fmt='import sys,json; d=json.loads(sys.stdin.read()); print(d["tag"], [round(s["us_per_step_best"],2) for s in d["steps"]])'
for t in base p1 base p1; do
DVG_LIB_TAG=$t timeout 200 python scripts/step_time.py --tag $t 2>&1 | tail -1 | python -c "$fmt"
done
for t in base p1; do
DVG_LIB_TAG=$t DVG_STEP_CHAIN=0 timeout 200 python scripts/step_time.py --tag ${t}nochain 2>&1 | tail -1 | python -c "$fmt"
done
DVG_LIB_TAG=p1 timeout 600 python -m pytest tests/test_gpu_rollout.py tests/test_gpu_gp.py -x -q -m gpu 2>&1 | tail -3
