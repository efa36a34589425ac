#!/usr/bin/env bash
# Round 2, GPU session D: tb2x ring-depth sweep after the early stage release.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 300 python tools/check_tb2x.py > $O/d_check_tb2x.log 2>&1; tail -1 $O/d_check_tb2x.log
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu"
for sd in "4 3 0" "5 3 0" "5 4 0" "6 3 0" "5 3 23040" "4 4 0" "3 3 0" "5 2 0"; do set -- $sd
  FDTD_B200_TB2X_STAGES=$1 FDTD_B200_TB2X_SLOTS=$2 FDTD_B200_TB2X_PAD=$3 timeout 300 $B > $O/d_bench_tb2x_s$1_d$2_p$3.json 2>&1
done
for f in $O/d_bench_*.json; do echo "== $f"; tail -1 $f | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print(round(d['value']/1e9,2),'Gcell/s frac',round(d['roofline']['frac'],3), d['clocks'])
except Exception as e: print('unparsed', e)"; done
