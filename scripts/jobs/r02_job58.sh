export DVG_TRACE=1 DVG_LIB_TAG=trace DVG_TC_TRACE=1
timeout 200 python scripts/chain_trace.py --kind plain --first 3 > /dev/null 2> gpurun_out/r02_chain_plain3.log
grep -c "^cta" gpurun_out/r02_chain_plain3.log
