timeout 700 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
export SWEEP_ROUNDS=2
export SWEEP_CFGS='[["default", null, {}], ["park 224", null, {"PARK_TICKS": 224}], ["park 192 b213", null, {"PARK_TICKS": 192, "SMEM_BUDGET_KB": 213}]]'
timeout 600 python tools/sweep_policy.py 2>&1 | tail -8
MIX=all timeout 120 python tools/trace_mix.py 2>&1 | tail -1
