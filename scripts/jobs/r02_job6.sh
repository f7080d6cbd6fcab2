#!/bin/bash
mkdir -p gpurun_out
for mega in 0 1; do
DVG_STEP_HEAD_MEGA=$mega python scripts/step_time.py --tag poll_mega$mega > gpurun_out/r02_steptime_poll_m$mega.json 2>/dev/null; cat gpurun_out/r02_steptime_poll_m$mega.json
DVG_STEP_HEAD_MEGA=$mega python scripts/step_time.py --workload bair_s32 --tag poll_mega$mega > gpurun_out/r02_steptime_poll_bair_m$mega.json 2>/dev/null; cat gpurun_out/r02_steptime_poll_bair_m$mega.json
done
DVG_STEP_HEAD_MEGA=0 DVG_TRACE=1 DVG_LIB_TAG=trace DVG_TC_TRACE=1 DVG_TC_TRACE_LAUNCH=20 timeout 300 python scripts/profile_step.py --steps 30 --workload bair_s32 > /dev/null 2> gpurun_out/r02_trace_bair2.log
( time timeout 900 python -m pytest tests/test_gpu_lstm.py tests/test_gpu_rollout.py -x -q -m gpu ) > gpurun_out/r02_t6.log 2>&1; tail -3 gpurun_out/r02_t6.log
