#!/usr/bin/env bash
# Round 2, GPU session S: suite after the fp64 het change, c3 fp64, final physics line, launch lists (step kernels only).
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 ) > $O/s_pytest_gpu.log 2>&1; echo "PYTEST: $(grep -E ' passed| failed' $O/s_pytest_gpu.log | tail -1)"
show() { for f in "$@"; do echo "== $f"; tail -1 $f | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print(round(d['value']/1e9,2),'Gcell/s frac',round(d['roofline']['frac'],3), 'kernel ms/step', round(d['roofline']['kernel_ms_per_step'],3), 'traffic', d['roofline'].get('traffic'), d['clocks'], d.get('check') and d['check'].get('ok'))
except Exception as e: print('unparsed', e)"; done; }
timeout 300 python bench.py --workload c3 --steps 20 --warmup 3 --no-cpu --no-e2e --dtype float64 > $O/s_bench_c3_f64.json 2>&1
timeout 300 python bench.py --workload 512x512x256 --het --steps 20 --warmup 3 --no-cpu --no-e2e --dtype float64 > $O/s_bench_c3het_only_f64.json 2>&1
timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --physics > $O/s_bench_yeex_f32.json 2>&1
timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --physics --dtype float64 > $O/s_bench_yeex_f64.json 2>&1
show $O/s_bench_*.json
K='regex:k_fused3d|k_sources|k_monitors|k_bump|k_ade|k_h3d|k_e3d'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 200 --csv --log-file $O/s_launches_c4_twostep.csv python bench.py --steps 8 --warmup 3 --no-cpu --no-e2e --no-check > /dev/null 2>&1; wc -l $O/s_launches_c4_twostep.csv
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 200 --csv --log-file $O/s_launches_c4_physics.csv python bench.py --steps 8 --warmup 3 --no-cpu --no-e2e --no-check --physics > /dev/null 2>&1; wc -l $O/s_launches_c4_physics.csv
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 200 --csv --log-file $O/s_launches_c3.csv python bench.py --workload c3 --steps 8 --warmup 3 --no-cpu --no-e2e --no-check > /dev/null 2>&1; wc -l $O/s_launches_c3.csv
