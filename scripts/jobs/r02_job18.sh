python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02_bench2.json 2> gpurun_out/r02_bench2.err; tail -2 gpurun_out/r02_bench2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench2.json').read().strip().splitlines()[-1])
print('N=1 value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'e2e ms',d['e2e']['ms_per_step'], 'h2d', d['e2e']['h2d_bytes_per_step'], 'fired', d['config']['triggered_rollout_steps'])
PY
python -m pytest tests/test_gpu_rollout.py -q -m gpu -k "pipeline or graph" 2>&1 | tail -2
