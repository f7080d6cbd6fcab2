#!/bin/bash
bash scripts/jobs/r02_benchall.sh
bash scripts/jobs/r02_ncu.sh
ncu -i gpurun_out/r02_step.ncu-rep --page raw --csv > gpurun_out/r02_lstm_step_ncu_raw.csv 2>/dev/null
ncu -i gpurun_out/r02_small.ncu-rep --page raw --csv > gpurun_out/r02_lstm_small_ncu_raw.csv 2>/dev/null
rm -f gpurun_out/r02_step.ncu-rep gpurun_out/r02_small.ncu-rep
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
