#!/usr/bin/env bash
# Round 2, GPU session G: physics sweep v2 (14/15 owner rows, psi loaded one iteration ahead, fp32 CPML math): tests, timing, ncu.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
( time timeout 900 python -m pytest tests/test_physics_mode.py -m gpu -q -x --timeout 600 ) > $O/g_pytest_physics.log 2>&1; tail -5 $O/g_pytest_physics.log
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --physics"
for pr in 3 0 1 2; do FDTD_B200_YEEX_L2PROMO=$pr timeout 300 $B > $O/g_bench_yeex_promo$pr.json 2>&1; done
for sd in "3 3" "5 3" "4 2"; do set -- $sd
  FDTD_B200_YEEX_STAGES=$1 FDTD_B200_YEEX_SLOTS=$2 timeout 300 $B > $O/g_bench_yeex_s$1_d$2.json 2>&1
done
timeout 300 $B --dtype float64 > $O/g_bench_yeex_f64.json 2>&1
for f in $O/g_bench_*.json; do echo "== $f"; tail -1 $f | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print(round(d['value']/1e9,2),'Gcell/s frac',round(d['roofline']['frac'],3), d['clocks'])
except Exception as e: print('unparsed', e)"; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused3d_yeex -s 2 -c 1 -o $O/g_yeex python bench.py --steps 4 --warmup 4 --no-cpu --no-e2e --physics > $O/g_ncu_yeex.log 2>&1; tail -2 $O/g_ncu_yeex.log
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_run.py yee > $O/g_sanitize_memcheck_yee.log 2>&1; tail -3 $O/g_sanitize_memcheck_yee.log
