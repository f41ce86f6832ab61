timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python tools/small_pop.py 2>&1 | tail -9
