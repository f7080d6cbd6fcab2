#!/bin/bash
mkdir -p gpurun_out
timeout 120 scripts/_bin/mma_probe 2000 0 > gpurun_out/r02_mma_probe.jsonl 2>&1
echo "probe rc=$?"
cat gpurun_out/r02_mma_probe.jsonl
timeout 120 scripts/_bin/mma_probe 1000 1 > gpurun_out/r02_mma_probe_direct.jsonl 2>&1
echo "probe-direct rc=$?"
tail -3 gpurun_out/r02_mma_probe_direct.jsonl
