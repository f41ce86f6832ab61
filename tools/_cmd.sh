timeout 700 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
echo "all classes"; MIX=all timeout 120 python tools/trace_mix.py 2>&1 | tail -1
echo "nb 2..8"; MIX=2..8 timeout 120 python tools/trace_mix.py 2>&1 | tail -1
SWEEP_ROUNDS=1 timeout 300 python tools/sweep_policy.py 2>&1 | tail -8
