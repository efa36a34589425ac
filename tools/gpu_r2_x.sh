#!/usr/bin/env bash
# Round 2, GPU session X: material-index coding in the heterogeneous sweep (one byte per cell + shared-memory material
# table instead of 4 / 6 coefficient arrays): equivalence tests, memcheck, full suite, c3 / c5 lines both ways, launch list.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
( time timeout 400 python -m pytest tests/test_raster.py tests/test_multi_slab.py -m gpu -q --timeout 300 -k "index_coded or indexed" ) > $O/x_pytest_new.log 2>&1
echo "NEW: $(grep -E ' passed| failed| error' $O/x_pytest_new.log | tail -1)"; grep -E "^(FAILED|ERROR)|^E  " $O/x_pytest_new.log | head -12
( timeout 200 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_raster.py -m gpu -q --timeout 180 -k "index_coded and float32" ) > $O/x_sanitize_memcheck_indexed.log 2>&1
echo "MEMCHECK rc=$? $(grep -E 'ERROR SUMMARY| passed| failed' $O/x_sanitize_memcheck_indexed.log | tail -2 | tr '\n' ' ')"
( time timeout 1200 python -m pytest tests -m gpu -q --timeout 900 ) > $O/x_pytest_gpu.log 2>&1; echo "PYTEST: $(grep -E ' passed| failed' $O/x_pytest_gpu.log | tail -1)"; grep -E "^(FAILED|ERROR)" $O/x_pytest_gpu.log | head
show() { for f in "$@"; do echo "== $f"; tail -1 $f | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print(round(d['value']/1e9,2),'Gcell/s frac',round(d['roofline']['frac'],3), 'kernel ms/step', round(d['roofline']['kernel_ms_per_step'],3), d['clocks']['sm_mhz'], d['clocks']['reasons'], d.get('check') and (d['check'].get('ok'), d['check'].get('crop_rel_l2'), d['check'].get('fields_sha')), d.get('setup'))
except Exception as e: print('unparsed', e)"; done; }
C="python bench.py --steps 40 --warmup 4 --no-cpu --no-e2e"
timeout 200 $C --workload c3 > $O/x_bench_c3_indexed.json 2>&1
timeout 200 $C --workload c3 --no-indexed --no-check > $O/x_bench_c3_arrays.json 2>&1
timeout 200 $C --workload c3 --aniso > $O/x_bench_c3_aniso_indexed.json 2>&1
show $O/x_bench_c3_indexed.json $O/x_bench_c3_arrays.json $O/x_bench_c3_aniso_indexed.json
timeout 400 python bench.py --workload c5 --steps 20 --warmup 3 --no-cpu --no-e2e > $O/x_bench_c5_aniso_indexed_1gpu.json 2>&1
show $O/x_bench_c5_aniso_indexed_1gpu.json
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_fused3d_het|k_rasterize" -c 8 --csv --log-file $O/x_launches_c3_aniso_indexed.csv python bench.py --workload c3 --aniso --steps 4 --warmup 3 --no-cpu --no-e2e --no-check > /dev/null 2>&1
grep -E "k_rasterize|k_fused3d_het" $O/x_launches_c3_aniso_indexed.csv | awk -F'","' '{print substr($5,1,44), $(NF-2), $(NF-1), $NF}' | head -12
