#!/bin/bash
# round 2, GPU job 1: tensor-pipe / ring probe, per-k-block trace of lstm_step_kernel, baseline bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_smi.txt 2>&1
timeout 120 scripts/_bin/mma_probe 2000 0 > gpurun_out/r02_mma_probe.jsonl 2>&1
echo "probe rc=$?"
DVG_TRACE=1 DVG_LIB_TAG=trace DVG_TC_TRACE=1 DVG_TC_TRACE_LAUNCH=20 timeout 300 python scripts/profile_step.py --steps 30 > gpurun_out/r02_trace0.out 2> gpurun_out/r02_trace0.log
echo "trace rc=$?"
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench0.json 2> gpurun_out/r02_bench0.err
echo "bench rc=$?"
tail -c 600 gpurun_out/r02_bench0.json
timeout 120 scripts/_bin/mma_probe 1000 1 > gpurun_out/r02_mma_probe_direct.jsonl 2>&1
echo "probe-direct rc=$?"
cat gpurun_out/r02_mma_probe.jsonl
tail -3 gpurun_out/r02_mma_probe_direct.jsonl
