#!/usr/bin/env bash
# Round 2, GPU session P: box-first dispatch of the recursion-carrying sweeps, physics sweep without the provider row's E stage.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_ade_fused.py tests/test_gpu_configs.py tests/test_physics_mode.py tests/test_gpu_parity.py -m gpu -q --timeout 600 ) > $O/p_pytest.log 2>&1; tail -4 $O/p_pytest.log
show() { for f in "$@"; do echo "== $f"; tail -1 $f | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print(round(d['value']/1e9,2),'Gcell/s frac',round(d['roofline']['frac'],3), 'kernel ms/step', round(d['roofline']['kernel_ms_per_step'],3), 'post', round(d['roofline']['post_ms_per_step'],3), d['clocks'], d.get('check') and d['check'].get('ok'))
except Exception as e: print('unparsed', e)"; done; }
C="python bench.py --workload c3 --steps 40 --warmup 4 --no-cpu --no-e2e"
timeout 300 $C > $O/p_bench_c3_f32.json 2>&1
FDTD_B200_ADE_FUSED=0 timeout 300 $C --no-check > $O/p_bench_c3_f32_postade.json 2>&1
timeout 600 python bench.py --workload c5 --steps 20 --warmup 3 --no-cpu --no-e2e > $O/p_bench_c5_1gpu.json 2>&1
show $O/p_bench_c3_f32.json $O/p_bench_c3_f32_postade.json $O/p_bench_c5_1gpu.json
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --physics"
timeout 300 $B > $O/p_bench_yeex_f32.json 2>&1
timeout 300 $B --dtype float64 > $O/p_bench_yeex_f64.json 2>&1
show $O/p_bench_yeex_f32.json $O/p_bench_yeex_f64.json
