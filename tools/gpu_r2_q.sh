#!/usr/bin/env bash
# Round 2, GPU session Q: whole suite after the recursion indexing fix, memcheck of the recursion-carrying sweeps, c3 / c5 lines.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 ) > $O/q_pytest_gpu.log 2>&1; echo "PYTEST: $(grep -E ' passed| failed' $O/q_pytest_gpu.log | tail -1)"
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_run.py het yee tb2 > $O/q_sanitize_memcheck.log 2>&1; echo "MEMCHECK: $(tail -2 $O/q_sanitize_memcheck.log | tr '\n' ' ')"
show() { for f in "$@"; do echo "== $f"; tail -1 $f | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print(round(d['value']/1e9,2),'Gcell/s frac',round(d['roofline']['frac'],3), 'kernel ms/step', round(d['roofline']['kernel_ms_per_step'],3), 'post', round(d['roofline']['post_ms_per_step'],3), d['clocks'], d.get('check') and d['check'].get('ok'))
except Exception as e: print('unparsed', e)"; done; }
C="python bench.py --workload c3 --steps 40 --warmup 4 --no-cpu --no-e2e"
timeout 300 $C > $O/q_bench_c3_f32.json 2>&1
timeout 300 $C --dtype float64 --steps 20 > $O/q_bench_c3_f64.json 2>&1
timeout 600 python bench.py --workload c5 --steps 20 --warmup 3 --no-cpu --no-e2e > $O/q_bench_c5_1gpu.json 2>&1
show $O/q_bench_c3_f32.json $O/q_bench_c3_f64.json $O/q_bench_c5_1gpu.json
