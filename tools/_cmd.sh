timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 300 python bench.py --steps 3 --warmup 3 2>gpurun_out/bench_final.err | tee gpurun_out/bench_final.json | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum,launch__grid_size,launch__shared_mem_per_block_dynamic --clock-control none -c 2000 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1; tail -c 300 gpurun_out/bench_under_ncu.log; wc -l gpurun_out/launches_final.csv
timeout 600 ncu --set full --import-source on --clock-control none -k regex:episode_kernel --launch-skip 7 --launch-count 1 -o /tmp/bulk python tools/profile_bulk.py > gpurun_out/ncu_bulk.log 2>&1; tail -3 gpurun_out/ncu_bulk.log
ncu -i /tmp/bulk.ncu-rep --page raw --csv > gpurun_out/bulk_raw.csv 2>/dev/null
ncu -i /tmp/bulk.ncu-rep --page source --csv > gpurun_out/bulk_src.csv 2>/dev/null
ls -la gpurun_out/bulk_*.csv
