timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 python bench.py 2>gpurun_out/bench_n1.err | tee gpurun_out/bench_r1_n1.json | cut -c1-200
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 2>/dev/null | tee gpurun_out/bench_r1_ref.json | cut -c1-200
