#!/usr/bin/env bash
# Round 2, GPU session M: physics sweep v3 (compile-time psi families), c3 slowdown diagnosis (in-sweep recursions vs k_ade).
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
( time timeout 900 python -m pytest tests/test_physics_mode.py -m gpu -q -x --timeout 600 ) > $O/m_pytest_physics.log 2>&1; tail -4 $O/m_pytest_physics.log
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --physics"
timeout 300 $B > $O/m_bench_yeex.json 2>&1
FDTD_B200_BENCH_CPML=0 timeout 300 $B --no-check > $O/m_bench_yeex_nocpml.json 2>&1
timeout 300 $B --dtype float64 > $O/m_bench_yeex_f64.json 2>&1
C="python bench.py --workload c3 --steps 40 --warmup 4 --no-cpu --no-e2e --no-check"
timeout 300 $C > $O/m_bench_c3_fused.json 2>&1
FDTD_B200_ADE_FUSED=0 timeout 300 $C > $O/m_bench_c3_postade.json 2>&1
for f in $O/m_bench_*.json; do echo "== $f"; tail -1 $f | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print(round(d['value']/1e9,2),'Gcell/s frac',round(d['roofline']['frac'],3), 'kernel ms/step', round(d['roofline']['kernel_ms_per_step'],3), 'post', round(d['roofline']['post_ms_per_step'],3), d['clocks'])
except Exception as e: print('unparsed', e)"; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused3d_het -s 3 -c 1 -o $O/m_het_ade $C > $O/m_ncu_het.log 2>&1; tail -2 $O/m_ncu_het.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused3d_yeex -s 2 -c 1 -o $O/m_yeex python bench.py --steps 4 --warmup 4 --no-cpu --no-e2e --physics --no-check > $O/m_ncu_yeex.log 2>&1; tail -2 $O/m_ncu_yeex.log
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_run.py yee het > $O/m_sanitize_memcheck.log 2>&1; tail -3 $O/m_sanitize_memcheck.log
