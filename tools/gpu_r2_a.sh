#!/usr/bin/env bash
# Round 2, GPU session A: full GPU test-suite, the bench lines VERDICT asked for (with clocks), sanitizer runs.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $O/a_smi.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 ) > $O/a_pytest.log 2>&1
tail -5 $O/a_pytest.log
B="python bench.py --steps 20 --warmup 3"
timeout 600 $B > $O/a_bench_c4_f32.json 2> $O/a_bench_c4_f32.err
timeout 600 $B --dtype float64 --no-e2e --no-cpu > $O/a_bench_c4_f64.json 2> $O/a_bench_c4_f64.err
FDTD_B200_EXACT_DIV=1 timeout 600 $B --dtype float64 --no-e2e --no-cpu --steps 6 > $O/a_bench_c4_f64_ddiv.json 2>&1
timeout 600 $B --dtype float64 --fast-f64 --no-e2e --no-cpu > $O/a_bench_c4_f64_fast.json 2>&1
FDTD_B200_TB2=0 timeout 600 $B --dtype float64 --no-e2e --no-cpu > $O/a_bench_c4_f64_onestep.json 2>&1
timeout 300 python bench.py --workload c1 --steps 4000 --warmup 64 > $O/a_bench_c1_f32.json 2>&1
timeout 300 python bench.py --workload c1 --steps 4000 --warmup 64 --dtype float64 > $O/a_bench_c1_f64.json 2>&1
timeout 300 python bench.py --workload c2 --steps 4000 --warmup 64 --no-e2e --no-cpu > $O/a_bench_c2_f32.json 2>&1
timeout 300 python bench.py --workload c3 --het --steps 40 --warmup 4 --no-e2e --no-cpu > $O/a_bench_c3het_f32.json 2>&1
timeout 600 $B --physics --no-e2e --no-cpu > $O/a_bench_c4_physics_twopass.json 2>&1
FDTD_B200_YEE_FUSED=1 timeout 600 $B --physics --no-e2e --no-cpu > $O/a_bench_c4_physics_fused.json 2>&1
timeout 600 python tools/check_yee_fused.py 512 > $O/a_yee_fused.log 2>&1
for t in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $t python tools/sanitize_run.py > $O/a_sanitize_$t.log 2>&1
  tail -3 $O/a_sanitize_$t.log
done
for f in $O/a_bench_*.json; do echo "== $f"; tail -c 1500 $f | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print(round(d['value']/1e9,2),'Gcell/s frac',round(d['roofline']['frac'],3), d['roofline'].get('kernel'), d['clocks'])
except Exception as e: print('unparsed', e)"; done
