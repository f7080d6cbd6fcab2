python bench.py --steps 10 --warmup 3 --cpu-budget 5 > gpurun_out/r02_bench1.json 2> gpurun_out/r02_bench1.err; tail -3 gpurun_out/r02_bench1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench1.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])
print('stock',json.dumps(d['stock_torch_b200'],indent=0))
print('pixel',json.dumps(d['pixel_e2e'],indent=0))
print('cpu',d['cpu_baseline'])
PY
