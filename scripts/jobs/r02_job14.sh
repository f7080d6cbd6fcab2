bash scripts/ab4.sh "DVG_STEP_WARM=0 DVG_LIB_TAG=warm0" "DVG_STEP_X=auxmask DVG_LIB_TAG=auxmask" "DVG_STEP_X=auxmask DVG_LIB_TAG=auxmask DVG_STEP_SCHED=2"
DVG_STEP_X=auxmask DVG_LIB_TAG=auxmask python -m pytest tests/test_gpu_lstm.py tests/test_gpu_rollout.py -x -q -m gpu 2>&1 | tail -3
DVG_STEP_X=auxmask DVG_LIB_TAG=auxmask python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['lstm_step_ms'])"
DVG_STEP_SCHED=2 DVG_STEP_X=auxmask DVG_LIB_TAG=auxmask python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench sched2', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['lstm_step_ms'])"
