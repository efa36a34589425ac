#!/usr/bin/env bash
# Round 2, GPU session U: physics sweep with 19 consumer rows (fp32, 640 threads, 96 registers).
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
( time timeout 900 python -m pytest tests/test_physics_mode.py -m gpu -q --timeout 600 ) > $O/u_pytest_physics.log 2>&1; echo "PYTEST: $(grep -E ' passed| failed' $O/u_pytest_physics.log | tail -1)"
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --physics --no-check"
for i in 1 2; do timeout 300 $B > $O/u_yeex_r19_$i.json 2>&1; done
for lx in 32 128; do FDTD_B200_FUSED_LX=$lx timeout 300 $B > $O/u_yeex_r19_lx$lx.json 2>&1; done
FDTD_B200_YEEX_STAGES=3 timeout 300 $B > $O/u_yeex_r19_s3.json 2>&1
for f in $O/u_yeex_*.json; do echo "== $f"; tail -1 $f | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print(round(d['value']/1e9,2),'Gcell/s frac',round(d['roofline']['frac'],3), 'kernel ms/step', round(d['roofline']['kernel_ms_per_step'],3), d['clocks'])
except Exception as e: print('unparsed', e)"; done
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_run.py yee > $O/u_sanitize_memcheck.log 2>&1; echo "MEMCHECK: $(tail -2 $O/u_sanitize_memcheck.log | tr '\n' ' ')"
