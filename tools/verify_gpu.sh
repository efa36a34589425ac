timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 3000 gpurun_out/bench_final.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-600
