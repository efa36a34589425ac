#!/usr/bin/env bash
# Round 2, GPU session V: device geometry rasterisation (row f4) + per-component Cb in the het sweep (row f3):
# new tests, memcheck on them, full suite, c3 / c5 lines with device-painted coefficients, launch list.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
( time timeout 600 python -m pytest tests/test_raster.py tests/test_gpu_plugin.py tests/test_multi_slab.py -m gpu -q --timeout 400 -k "raster or geometry or aniso or per_component or equal_components or long_shape or session_set" ) > $O/v_pytest_new.log 2>&1
echo "NEW: $(grep -E ' passed| failed| error' $O/v_pytest_new.log | tail -1)"; grep -E "^(FAILED|ERROR)|Error|assert" $O/v_pytest_new.log | head -20
( timeout 400 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_raster.py -m gpu -q --timeout 380 -k "oracle_and_reference or per_component or long_shape" ) > $O/v_sanitize_memcheck_raster.log 2>&1
echo "MEMCHECK rc=$? $(grep -E 'ERROR SUMMARY| passed| failed' $O/v_sanitize_memcheck_raster.log | tail -2 | tr '\n' ' ')"
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 ) > $O/v_pytest_gpu.log 2>&1; echo "PYTEST: $(grep -E ' passed| failed' $O/v_pytest_gpu.log | tail -1)"
show() { for f in "$@"; do echo "== $f"; tail -1 $f | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print(round(d['value']/1e9,2),'Gcell/s frac',round(d['roofline']['frac'],3), 'kernel ms/step', round(d['roofline']['kernel_ms_per_step'],3), d['clocks']['sm_mhz'], d['clocks']['reasons'], d.get('check') and (d['check'].get('ok'), d['check'].get('crop_rel_l2')), d.get('setup'))
except Exception as e: print('unparsed', e)"; done; }
C="python bench.py --steps 40 --warmup 4 --no-cpu --no-e2e"
timeout 300 $C --workload c3 > $O/v_bench_c3_raster.json 2>&1
timeout 300 $C --workload c3 --host-coeffs --no-check > $O/v_bench_c3_hostcoeffs.json 2>&1
timeout 300 $C --workload c3 --aniso > $O/v_bench_c3_aniso.json 2>&1
show $O/v_bench_c3_raster.json $O/v_bench_c3_hostcoeffs.json $O/v_bench_c3_aniso.json
timeout 500 python bench.py --workload c5 --steps 20 --warmup 3 --no-cpu --no-e2e > $O/v_bench_c5_aniso_1gpu.json 2>&1
timeout 500 python bench.py --workload c5 --no-aniso --steps 20 --warmup 3 --no-cpu --no-e2e --no-check > $O/v_bench_c5_iso_1gpu.json 2>&1
show $O/v_bench_c5_aniso_1gpu.json $O/v_bench_c5_iso_1gpu.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/v_launches_c3_aniso.csv python bench.py --workload c3 --aniso --steps 4 --warmup 3 --no-cpu --no-e2e --no-check > $O/v_ncu_c3_aniso.log 2>&1
grep -E "k_rasterize|k_fused3d_het" $O/v_launches_c3_aniso.csv | awk -F'","' '{print $5, $NF}' | sort | uniq -c | sort -rn | head -8
