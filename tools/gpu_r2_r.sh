#!/usr/bin/env bash
# Round 2, GPU session R: evict-first store experiment on the physics sweep, final ncu captures, the driver's default lines.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
show() { for f in "$@"; do echo "== $f"; tail -1 $f | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print(round(d['value']/1e9,2),'Gcell/s frac',round(d['roofline']['frac'],3), 'kernel ms/step', round(d['roofline']['kernel_ms_per_step'],3), 'e2e', d['e2e'] and round(d['e2e']['value']/1e9,2), d['clocks'], d.get('check') and d['check'].get('ok'))
except Exception as e: print('unparsed', e)"; done; }
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --physics --no-check"
for cs in 0 1 0 1; do FDTD_B200_YEEX_STCS=$cs timeout 300 $B > $O/r_yeex_stcs${cs}_$RANDOM.json 2>&1; done
show $O/r_yeex_stcs*.json
( time python bench.py ) > $O/r_bench_default.json 2> $O/r_bench_default.err; show $O/r_bench_default.json; tail -4 $O/r_bench_default.err
( time python bench.py --impl reference ) > $O/r_bench_reference.json 2> $O/r_bench_reference.err; tail -c 400 $O/r_bench_reference.json; tail -4 $O/r_bench_reference.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/r_smoke.log 2>&1; tail -3 $O/r_smoke.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused3d_yeex -s 2 -c 1 -o $O/r_yeex python bench.py --steps 4 --warmup 4 --no-cpu --no-e2e --physics --no-check > $O/r_ncu_yeex.log 2>&1; tail -1 $O/r_ncu_yeex.log | cut -c1-100
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused3d_tb2x -s 2 -c 1 -o $O/r_tb2x python bench.py --steps 8 --warmup 4 --no-cpu --no-e2e --no-check > $O/r_ncu_tb2x.log 2>&1; tail -1 $O/r_ncu_tb2x.log | cut -c1-100
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused3d_het -s 3 -c 1 -o $O/r_het python bench.py --workload c3 --steps 8 --warmup 4 --no-cpu --no-e2e --no-check > $O/r_ncu_het.log 2>&1; tail -1 $O/r_ncu_het.log | cut -c1-100
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r_launches_c4.csv python bench.py --steps 8 --warmup 3 --no-cpu --no-e2e --no-check > $O/r_launches.log 2>&1; wc -l $O/r_launches_c4.csv
