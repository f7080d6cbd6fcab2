#!/bin/bash
mkdir -p gpurun_out
python scripts/step_time.py --tag mega > gpurun_out/r02_steptime_mega.json 2> gpurun_out/r02_steptime_mega.err; cat gpurun_out/r02_steptime_mega.json
python scripts/step_time.py --workload bair_s32 --tag mega > gpurun_out/r02_steptime_mega_bair.json 2>/dev/null; cat gpurun_out/r02_steptime_mega_bair.json
DVG_STEP_HEAD_MEGA=0 python scripts/step_time.py --tag nomega > gpurun_out/r02_steptime_nomega.json 2>/dev/null; cat gpurun_out/r02_steptime_nomega.json
( time timeout 900 python -m pytest tests/test_gpu_lstm.py tests/test_gpu_rollout.py -x -q -m gpu ) > gpurun_out/r02_t4.log 2>&1; tail -5 gpurun_out/r02_t4.log
