#!/usr/bin/env bash
# Round 2, GPU session K: whole GPU suite (device mode overlap included) and the config lines c1 / c2 / c3 / c4 with clocks.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 ) > $O/k_pytest_gpu.log 2>&1; tail -6 $O/k_pytest_gpu.log
run() { name=$1; shift; timeout 600 python bench.py "$@" > $O/k_bench_$name.json 2> $O/k_bench_$name.err; tail -1 $O/k_bench_$name.json | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('$name', round(d['value']/1e9,2),'Gcell/s frac',round(d['roofline']['frac'],3), 'e2e', d['e2e'] and round(d['e2e']['value']/1e9,2), 'check', d.get('check') and {k:d['check'].get(k) for k in ('crop_rel_l2','dft_rel_l2','ok','crop_bit_exact')}, d['clocks'])
except Exception as e: print('$name unparsed', e)"; tail -2 $O/k_bench_$name.err; }
run c3_f32 --workload c3 --steps 40 --warmup 4 --no-cpu
run c3_f64 --workload c3 --steps 20 --warmup 3 --no-cpu --no-e2e --dtype float64
run c3het_only_f32 --workload 512x512x256 --het --steps 40 --warmup 4 --no-cpu --no-e2e
run c2_f32 --workload c2 --steps 200 --warmup 10 --no-cpu --no-e2e
run c2_f64 --workload c2 --steps 200 --warmup 10 --no-cpu --no-e2e --dtype float64
run c1_f32 --workload c1 --steps 400 --warmup 20
run c1_f64 --workload c1 --steps 400 --warmup 20 --dtype float64
run c4_f64 --steps 20 --warmup 3 --no-cpu --no-e2e --dtype float64
run c4_f32_full --steps 60 --warmup 5
