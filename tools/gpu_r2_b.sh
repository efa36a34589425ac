#!/usr/bin/env bash
# Round 2, GPU session B: TMA two-step sweep — correctness first (under a short timeout), then A/B timing.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/dp_rate tools/dp_rate.cu && /tmp/dp_rate > $O/b_dp_rate.txt 2>&1; cat $O/b_dp_rate.txt
timeout 600 python tools/check_tb2x.py > $O/b_check_tb2x.log 2>&1; echo "check_tb2x rc=$?"; tail -12 $O/b_check_tb2x.log
if grep -q "TB2X_CHECK OK" $O/b_check_tb2x.log; then
  ( time timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 ) > $O/b_pytest.log 2>&1; tail -4 $O/b_pytest.log
fi
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu"
FDTD_B200_TB2X=0 timeout 300 $B > $O/b_bench_tb2.json 2>&1
for sd in "4 4" "5 3" "3 4" "6 2" "4 3"; do set -- $sd
  FDTD_B200_TB2X=1 FDTD_B200_TB2X_STAGES=$1 FDTD_B200_TB2X_SLOTS=$2 timeout 300 $B > $O/b_bench_tb2x_s$1_d$2.json 2>&1
done
FDTD_B200_TB2X=1 timeout 300 $B --no-ops > $O/b_bench_tb2x_noops.json 2>&1
FDTD_B200_TB2X=1 timeout 300 $B --dtype float64 > $O/b_bench_tb2x_f64.json 2>&1
for f in $O/b_bench_*.json; do echo "== $f"; tail -1 $f | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print(round(d['value']/1e9,2),'Gcell/s frac',round(d['roofline']['frac'],3), d['clocks'])
except Exception as e: print('unparsed', e)"; done
for t in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $t python tools/sanitize_run.py tb2 > $O/b_sanitize_$t.log 2>&1
  tail -3 $O/b_sanitize_$t.log
done
