timeout 400 python -m pytest tests/test_gpu_lstm.py tests/test_gpu_rollout.py -x -q -m gpu 2>&1 | tail -4
bash scripts/ab4.sh "DVG_LIB_TAG=wide DVG_STEP_X=wide DVG_STEP_HEAD_WIDE=0" "DVG_LIB_TAG=wide DVG_STEP_X=wide DVG_STEP_HEAD_WIDE=1"
WL=bair_s32 bash scripts/ab4.sh "DVG_LIB_TAG=wide DVG_STEP_X=wide DVG_STEP_HEAD_WIDE=0" "DVG_LIB_TAG=wide DVG_STEP_X=wide DVG_STEP_HEAD_WIDE=1"
