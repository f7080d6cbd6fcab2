export DVG_TRACE=1 DVG_LIB_TAG=trace DVG_TC_TRACE=1
for k in warm decide; do
timeout 200 python scripts/chain_trace.py --kind $k > /dev/null 2> gpurun_out/r02_chain2_$k.log
grep -c "^cta" gpurun_out/r02_chain2_$k.log
done
