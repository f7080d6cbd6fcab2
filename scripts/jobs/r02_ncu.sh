#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 300 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r02_ncu_b.log 2>&1
echo "launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_step_kernel -s 20 -c 1 -f -o gpurun_out/r02_step python scripts/profile_step.py --steps 30 > gpurun_out/r02_ncu_step.log 2>&1
echo "step rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lstm_small_kernel -s 5 -c 1 -f -o gpurun_out/r02_small python scripts/small_trace.py > gpurun_out/r02_ncu_small.log 2>&1
echo "small rc=$?"
ls -la gpurun_out/r02_step.ncu-rep gpurun_out/r02_small.ncu-rep gpurun_out/r02_launches.csv
