export DVG_LIB_TAG=rt1
timeout 300 python -c "import dvg_b200.build as b; b.build()" 2>&1 | tail -2
DVG_TC_SMALL=0 timeout 600 python -m pytest tests/test_gpu_lstm.py -x -q -m gpu 2>&1 | tail -5
timeout 600 python -m pytest tests/test_gpu_lstm.py tests/test_gpu_rollout.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python scripts/hidden_sweep.py 2>/dev/null
