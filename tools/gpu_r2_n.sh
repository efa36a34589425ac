#!/usr/bin/env bash
# Round 2, GPU session N: c3 with CTA-level recursion dispatch, physics sweep tuning (segment length, ring depths), suite.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 ) > $O/n_pytest_gpu.log 2>&1; tail -5 $O/n_pytest_gpu.log
show() { for f in "$@"; do echo "== $f"; tail -1 $f | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print(round(d['value']/1e9,2),'Gcell/s frac',round(d['roofline']['frac'],3), 'kernel ms/step', round(d['roofline']['kernel_ms_per_step'],3), 'post', round(d['roofline']['post_ms_per_step'],3), d['clocks'], d.get('check') and d['check'].get('ok'))
except Exception as e: print('unparsed', e)"; done; }
C="python bench.py --workload c3 --steps 40 --warmup 4 --no-cpu --no-e2e"
timeout 300 $C > $O/n_bench_c3_f32.json 2>&1
timeout 300 $C --dtype float64 --steps 20 > $O/n_bench_c3_f64.json 2>&1
timeout 600 python bench.py --workload c5 --steps 20 --warmup 3 --no-cpu --no-e2e > $O/n_bench_c5_1gpu.json 2>&1
show $O/n_bench_c3_f32.json $O/n_bench_c3_f64.json $O/n_bench_c5_1gpu.json
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --physics --no-check"
for lx in 64 128 342 512 1024; do FDTD_B200_FUSED_LX=$lx timeout 300 $B > $O/n_yeex_lx$lx.json 2>&1; done
for sd in "3 3" "5 4" "4 5" "6 3"; do set -- $sd
  FDTD_B200_YEEX_STAGES=$1 FDTD_B200_YEEX_SLOTS=$2 timeout 300 $B > $O/n_yeex_s$1_d$2.json 2>&1
done
show $O/n_yeex_*.json
