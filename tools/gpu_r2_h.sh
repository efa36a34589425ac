#!/usr/bin/env bash
# Round 2, GPU session H: in-sweep ADE (deferred + coupled), self-checking bench, whole GPU suite, physics lean ceiling.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_ade_fused.py -m gpu -q -x --timeout 600 ) > $O/h_pytest_ade.log 2>&1; tail -12 $O/h_pytest_ade.log
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 ) > $O/h_pytest_gpu.log 2>&1; tail -8 $O/h_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > $O/h_bench_c4_check.json 2> $O/h_bench_c4_check.err; tail -c 900 $O/h_bench_c4_check.json; tail -3 $O/h_bench_c4_check.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --dtype float64 --workload 512x512x512 > $O/h_bench_512_f64_check.json 2>&1; tail -c 600 $O/h_bench_512_f64_check.json
FDTD_B200_BENCH_CPML=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --physics --no-check > $O/h_bench_yeex_nocpml.json 2>&1
tail -1 $O/h_bench_yeex_nocpml.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('yeex without CPML', round(d['value']/1e9,2),'Gcell/s frac',round(d['roofline']['frac'],3))"
