DVG_TRACE=1 DVG_LIB_TAG=trace DVG_TC_TRACE=1 DVG_TC_TRACE_LAUNCH=20 timeout 300 python scripts/profile_step.py --steps 30 --workload bair_s32 > /dev/null 2> gpurun_out/r02_trace_bair5.log
echo ok
