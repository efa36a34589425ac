#!/usr/bin/env bash
# Round 2, GPU session W: suite after the raster kernel's per-material table + the CustomWaveform / MagneticDipole
# scenario; memcheck; ncu --set full of k_rasterize and the per-component-Cb het sweep; c4 default line; launch list.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 ) > $O/w_pytest_gpu.log 2>&1; echo "PYTEST: $(grep -E ' passed| failed' $O/w_pytest_gpu.log | tail -1)"; grep -E "^(FAILED|ERROR)" $O/w_pytest_gpu.log | head
( timeout 300 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_raster.py -m gpu -q --timeout 280 -k "oracle_and_reference or per_component or long_shape or each_shape" ) > $O/w_sanitize_memcheck_raster.log 2>&1
echo "MEMCHECK rc=$? $(grep -E 'ERROR SUMMARY| passed| failed' $O/w_sanitize_memcheck_raster.log | tail -2 | tr '\n' ' ')"
C3="python bench.py --workload c3 --aniso --steps 4 --warmup 3 --no-cpu --no-e2e --no-check"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_rasterize|k_fused3d_het" -c 5 -f -o $O/w_het_aniso $C3 > $O/w_ncu_full.log 2>&1; ls -la $O/w_het_aniso.ncu-rep
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/w_launches_c3_aniso.csv $C3 > /dev/null 2>&1
grep -E "k_rasterize|k_fused3d_het" $O/w_launches_c3_aniso.csv | awk -F'","' '{print substr($5,1,40), $NF}' | head -4
show() { for f in "$@"; do echo "== $f"; tail -1 $f | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print(round(d['value']/1e9,2),'Gcell/s frac',round(d['roofline']['frac'],3), 'kernel ms/step', round(d['roofline']['kernel_ms_per_step'],3), d['clocks']['sm_mhz'], d['clocks']['reasons'], d.get('check') and (d['check'].get('ok'), d['check'].get('crop_rel_l2'), d['check'].get('fields_sha')), d.get('setup'), 'e2e', d.get('e2e') and round(d['e2e']['value']/1e9,2))
except Exception as e: print('unparsed', e)"; done; }
timeout 300 python bench.py --workload c3 --steps 40 --warmup 4 --no-cpu --no-e2e > $O/w_bench_c3.json 2>&1
timeout 300 python bench.py --workload c3 --aniso --steps 40 --warmup 4 --no-cpu --no-e2e > $O/w_bench_c3_aniso.json 2>&1
timeout 400 python bench.py --no-cpu > $O/w_bench_c4.json 2> $O/w_bench_c4.err
show $O/w_bench_c3.json $O/w_bench_c3_aniso.json $O/w_bench_c4.json
