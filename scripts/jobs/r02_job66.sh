timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
timeout 600 python bench.py > gpurun_out/r02_bench_chain.json 2> gpurun_out/r02_bench_chain.err; tail -c 600 gpurun_out/r02_bench_chain.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_chain.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["value"], d["roofline"]["lstm_step_ms"], d["roofline"]["frac"], d["stock_torch_b200"]["ours_us_per_step"], d["stock_torch_b200"].get("ours_us_per_step_stream_ordered"), d["clocks"])
PY
