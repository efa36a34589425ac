#!/usr/bin/env bash
# Round 2, GPU session O: c3 with batched in-sweep recursions; physics sweep segment / ring fine tuning.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_ade_fused.py tests/test_gpu_configs.py tests/test_physics_mode.py -m gpu -q --timeout 600 ) > $O/o_pytest.log 2>&1; tail -4 $O/o_pytest.log
show() { for f in "$@"; do echo "== $f"; tail -1 $f | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print(round(d['value']/1e9,2),'Gcell/s frac',round(d['roofline']['frac'],3), 'kernel ms/step', round(d['roofline']['kernel_ms_per_step'],3), 'post', round(d['roofline']['post_ms_per_step'],3), d['clocks'], d.get('check') and d['check'].get('ok'))
except Exception as e: print('unparsed', e)"; done; }
C="python bench.py --workload c3 --steps 40 --warmup 4 --no-cpu --no-e2e"
timeout 300 $C > $O/o_bench_c3_f32.json 2>&1
timeout 300 $C --dtype float64 --steps 20 > $O/o_bench_c3_f64.json 2>&1
FDTD_B200_ADE_FUSED=0 timeout 300 $C --dtype float64 --steps 20 --no-check > $O/o_bench_c3_f64_postade.json 2>&1
timeout 600 python bench.py --workload c5 --steps 20 --warmup 3 --no-cpu --no-e2e > $O/o_bench_c5_1gpu.json 2>&1
FDTD_B200_ADE_FUSED=0 timeout 600 python bench.py --workload c5 --steps 20 --warmup 3 --no-cpu --no-e2e --no-check > $O/o_bench_c5_1gpu_postade.json 2>&1
show $O/o_bench_c3_f32.json $O/o_bench_c3_f64.json $O/o_bench_c3_f64_postade.json $O/o_bench_c5_1gpu.json $O/o_bench_c5_1gpu_postade.json
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --physics --no-check"
for lx in 32 48 64 96; do for d in 3 5; do FDTD_B200_FUSED_LX=$lx FDTD_B200_YEEX_SLOTS=$d timeout 300 $B > $O/o_yeex_lx${lx}_d$d.json 2>&1; done; done
FDTD_B200_YEEX_STAGES=3 FDTD_B200_YEEX_SLOTS=5 timeout 300 $B > $O/o_yeex_s3_d5.json 2>&1
show $O/o_yeex_*.json
