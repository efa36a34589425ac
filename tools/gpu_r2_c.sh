#!/usr/bin/env bash
# Round 2, GPU session C: tb2x after the tiling fix + branch-free exact fp64; ncu capture of the TMA sweep.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 600 python tools/check_tb2x.py > $O/c_check_tb2x.log 2>&1; echo "check_tb2x rc=$?"; tail -3 $O/c_check_tb2x.log
( time timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 ) > $O/c_pytest.log 2>&1; tail -4 $O/c_pytest.log
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu"
for sd in "5 3" "6 3" "6 2" "7 2"; do set -- $sd
  FDTD_B200_TB2X_STAGES=$1 FDTD_B200_TB2X_SLOTS=$2 timeout 300 $B > $O/c_bench_tb2x_s$1_d$2.json 2>&1
done
timeout 300 $B --dtype float64 > $O/c_bench_tb2x_f64.json 2>&1
FDTD_B200_TB2X=0 timeout 300 $B --dtype float64 > $O/c_bench_tb2_f64.json 2>&1
FDTD_B200_TB2=0 timeout 300 $B --dtype float64 > $O/c_bench_onestep_f64.json 2>&1
timeout 300 $B --dtype float64 --fast-f64 > $O/c_bench_tb2x_f64_fast.json 2>&1
FDTD_B200_TB2=0 timeout 300 $B > $O/c_bench_onestep_f32.json 2>&1
for f in $O/c_bench_*.json; do echo "== $f"; tail -1 $f | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print(round(d['value']/1e9,2),'Gcell/s frac',round(d['roofline']['frac'],3), d['clocks'])
except Exception as e: print('unparsed', e)"; done
FDTD_B200_TB2X_ARRIVE_ALL=1 timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_run.py tb2 > $O/c_sanitize_racecheck_tb2x_allarrive.log 2>&1; tail -3 $O/c_sanitize_racecheck_tb2x_allarrive.log
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_run.py > $O/c_sanitize_memcheck.log 2>&1; tail -3 $O/c_sanitize_memcheck.log
FDTD_B200_TB2X=0 timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_run.py > $O/c_sanitize_racecheck_all_tb2.log 2>&1; tail -3 $O/c_sanitize_racecheck_all_tb2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused3d_tb2x -s 2 -c 1 -o $O/c_tb2x python bench.py --steps 4 --warmup 4 --no-cpu --no-e2e > $O/c_ncu_tb2x.log 2>&1; tail -2 $O/c_ncu_tb2x.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused3d_yee -s 2 -c 1 -o $O/c_yee env FDTD_B200_YEE_FUSED=1 python bench.py --steps 4 --warmup 4 --no-cpu --no-e2e --physics > $O/c_ncu_yee.log 2>&1; tail -2 $O/c_ncu_yee.log
