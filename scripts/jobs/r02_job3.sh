#!/bin/bash
mkdir -p gpurun_out
for m in 1 2; do
DVG_STEP_SCHED=$m DVG_STEP_SCHED_VERBOSE=1 python scripts/step_time.py --tag sched$m > gpurun_out/r02_steptime_sched$m.json 2> gpurun_out/r02_steptime_sched$m.err; cat gpurun_out/r02_steptime_sched$m.json; grep schedule gpurun_out/r02_steptime_sched$m.err | head -2
done
for wl in bair_s32 ucf_s100; do
python scripts/step_time.py --workload $wl --tag base > gpurun_out/r02_steptime_$wl.json 2>/dev/null; cat gpurun_out/r02_steptime_$wl.json
done
