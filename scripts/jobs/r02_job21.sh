timeout 300 python -m pytest tests/test_gpu_lstm.py -x -q -m gpu 2>&1 | tail -8
for wl in smmnist_b16; do
DVG_TC_SMALL=0 timeout 120 python scripts/step_time.py --workload $wl --tag per_gemm 2>/dev/null | tail -1
timeout 120 python scripts/step_time.py --workload $wl --tag small 2>/dev/null | tail -1
done
