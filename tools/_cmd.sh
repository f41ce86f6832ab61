export SWEEP_ROUNDS=1
export SWEEP_CFGS='[["default", null, {}], ["small 0.8", null, {"SMALL_WEIGHT": 0.8}], ["small 0.65", null, {"SMALL_WEIGHT": 0.65}], ["small 0.5", null, {"SMALL_WEIGHT": 0.5}], ["small 0.65 park 224", null, {"SMALL_WEIGHT": 0.65, "PARK_TICKS": 224}], ["small 0.65 park 224 drain 8 b213", null, {"SMALL_WEIGHT": 0.65, "PARK_TICKS": 224, "DRAIN_LANES": 8, "SMEM_BUDGET_KB": 213}], ["default", null, {}]]'
timeout 600 python tools/sweep_policy.py 2>&1 | tail -9
