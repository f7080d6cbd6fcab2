#!/bin/bash
mkdir -p gpurun_out
python scripts/step_time.py --tag base > gpurun_out/r02_steptime_base.json 2> gpurun_out/r02_steptime_base.err; cat gpurun_out/r02_steptime_base.json
DVG_STEP_TRIG_EARLY=0 DVG_LIB_TAG=trig0 python scripts/step_time.py --tag trig_early0 > gpurun_out/r02_steptime_trig0.json 2> gpurun_out/r02_steptime_trig0.err; cat gpurun_out/r02_steptime_trig0.json
DVG_STEP_TRIG_EARLY=0 DVG_LIB_TAG=trig0 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_trig0.json 2> gpurun_out/r02_bench_trig0.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_trig0.json').read().strip().splitlines()[-1])
print('trig0 bench', d['value'], d['ms_per_step'], d['roofline']['lstm_step_ms'])
PY
