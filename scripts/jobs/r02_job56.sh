timeout 600 python -m pytest tests/test_gpu_rollout.py -x -q -m gpu -k "chained" 2>&1 | tail -5
timeout 300 python scripts/chain_debug.py 2>&1 | grep -E "^run|final" | head -12
export DVG_LIB_NOREBUILD=1
fmt='import sys,json; d=json.loads(sys.stdin.read()); print(d["tag"], [round(s["us_per_step_best"],2) for s in d["steps"]])'
for t in v5 v6; do
DVG_LIB_TAG=$t timeout 200 python scripts/step_time.py --tag $t 2>&1 | tail -1 | python -c "$fmt"
done
DVG_LIB_TAG=v6 DVG_STEP_CHAIN=0 timeout 200 python scripts/step_time.py --tag v6nochain 2>&1 | tail -1 | python -c "$fmt"
