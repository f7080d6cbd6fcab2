timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --no-extras > gpurun_out/r02_bench_check.json 2>/dev/null; echo "rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_check.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["lstm_step_ms"], d["clocks"])
PY
