timeout 600 python -m pytest tests/test_gpu_rollout.py -x -q -m gpu 2>&1 | tail -3
export DVG_LIB_NOREBUILD=1
fmt='import sys,json; d=json.loads(sys.stdin.read()); print(d["tag"], [round(s["us_per_step_best"],2) for s in d["steps"]])'
for t in v5 v11 v5 v11; do
DVG_LIB_TAG=$t timeout 200 python scripts/step_time.py --tag $t 2>&1 | tail -1 | python -c "$fmt"
done
DVG_LIB_TAG=v11 DVG_STEP_CHAIN=0 timeout 200 python scripts/step_time.py --tag v11nochain 2>&1 | tail -1 | python -c "$fmt"
DVG_LIB_TAG=v5 DVG_STEP_CHAIN=0 timeout 200 python scripts/step_time.py --tag v5nochain 2>&1 | tail -1 | python -c "$fmt"
DVG_LIB_TAG=v11 timeout 200 python scripts/step_time.py --workload bair_s32 --tag v11bair 2>&1 | tail -1 | python -c "$fmt"
