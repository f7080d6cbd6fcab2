#!/bin/bash
DVG_STEP_HEAD_MEGA=0 DVG_TRACE=1 DVG_LIB_TAG=trace DVG_TC_TRACE=1 DVG_TC_TRACE_LAUNCH=20 timeout 300 python scripts/profile_step.py --steps 30 --workload bair_s32 > /dev/null 2> gpurun_out/r02_trace_bair3.log
( time timeout 1500 python -m pytest tests -q -m gpu ) > gpurun_out/r02_t8.log 2>&1; tail -30 gpurun_out/r02_t8.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
