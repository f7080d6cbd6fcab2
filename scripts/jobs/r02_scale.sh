#!/bin/bash
# usage: r02_scale.sh N  -- one bench run at N GPUs (torchrun), output to gpurun_out/r02_scale_nN.json
N=$1
nvidia-smi topo -m > gpurun_out/r02_topo_n$N.txt 2>&1
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02_scale_n$N.json 2> gpurun_out/r02_scale_n$N.err
echo "rc=$?"; tail -3 gpurun_out/r02_scale_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_scale_n$N.json').read().strip().splitlines()[-1])
print('N=$N value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'e2e ms',d['e2e']['ms_per_step'], d['config'].get('multi_gpu'))
PY
