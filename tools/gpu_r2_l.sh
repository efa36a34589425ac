#!/usr/bin/env bash
# Round 2, GPU session L: c3 with the per-thread recursion mask, staged host<->device copies (e2e), c5 on one GPU.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_ade_fused.py tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q --timeout 600 ) > $O/l_pytest.log 2>&1; tail -4 $O/l_pytest.log
run() { name=$1; shift; timeout 900 python bench.py "$@" > $O/l_bench_$name.json 2> $O/l_bench_$name.err; tail -1 $O/l_bench_$name.json | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('$name', round(d['value']/1e9,2),'Gcell/s frac',round(d['roofline']['frac'],3), 'e2e', d['e2e'] and (round(d['e2e']['value']/1e9,2), round(d['e2e']['seconds'],2)), 'check', d.get('check') and {k:d['check'].get(k) for k in ('crop_rel_l2','dft_rel_l2','ok','crop_bit_exact')}, d['clocks'], d.get('s_params') and d['s_params']['S21'][:1])
except Exception as e: print('$name unparsed', e)"; tail -2 $O/l_bench_$name.err; }
run c3_f32 --workload c3 --steps 40 --warmup 4 --no-cpu
run c3_f64 --workload c3 --steps 20 --warmup 3 --no-cpu --no-e2e --dtype float64
run c4_e2e20 --steps 20 --warmup 3 --no-cpu
FDTD_B200_STAGED_COPY=0 run c4_e2e20_memcpy3d --steps 20 --warmup 3 --no-cpu --no-check
run c5_1gpu_f32 --workload c5 --steps 20 --warmup 3 --no-cpu --no-e2e
