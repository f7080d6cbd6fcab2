#!/bin/bash
# A/B/C/D of developer builds on one box; each argument is an env string
wl=${WL:-kth_s100}
for rep in 1 2; do
  for v in "$@"; do
    env $v DVG_STEP_HEAD_MEGA=${MEGA:-0} python scripts/step_time.py --workload $wl --tag "$v" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['tag'], d['workload'], ' '.join('%s=%.2f' % (s['kind'].split('_')[0]+'_'+s['kind'].split('_')[1], s['us_per_step_best']) for s in d['steps']))"
  done
done
