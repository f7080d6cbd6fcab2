#!/bin/bash
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "kth rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null; echo "ref rc=$?"
python bench.py --variant bf16 --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02_bench_bf16.json 2>/dev/null
for wl in bair_s32 ucf_s100 smmnist_b16; do
python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_wl_$wl.json 2> gpurun_out/r02_wl_$wl.err; echo "$wl rc=$?"
done
python - <<'PY'
import json
for f in ['r02_bench','r02_bench_bf16','r02_wl_bair_s32','r02_wl_ucf_s100','r02_wl_smmnist_b16']:
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f,'value %.4g'%d['value'],'ms %.4f'%d['ms_per_step'],'e2e %.4g'%d['e2e']['value'],'step_us %.2f'%(d['roofline']['lstm_step_ms']*1e3),'frac %.3f'%d['roofline']['frac'],'fired',d['config']['triggered_rollout_steps'], 'stock', (d.get('stock_torch_b200') or {}).get('graph_us_per_step_fp32'), (d.get('stock_torch_b200') or {}).get('ours_us_per_step'), 'pix', (d.get('pixel_e2e') or {}).get('value'))
    except Exception as e: print(f,'ERR',e)
PY
