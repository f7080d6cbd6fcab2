timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["lstm_step_ms"], d["roofline"]["lstm_step_ms_stream_ordered"], d["roofline"]["frac"], d["config"]["step_launches"][:40], d["clocks"], d["gpu_launches"])
PY
