#!/usr/bin/env bash
# Round 2, GPU session T: suite after the DFT-accumulator prefetch in the two-step sweep; c4 line.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 ) > $O/t_pytest_gpu.log 2>&1; echo "PYTEST: $(grep -E ' passed| failed' $O/t_pytest_gpu.log | tail -1)"
for i in 1 2; do timeout 300 python bench.py --steps 40 --warmup 4 --no-e2e --no-cpu > $O/t_bench_c4_$i.json 2>&1; tail -1 $O/t_bench_c4_$i.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['value']/1e9,2),'Gcell/s frac',round(d['roofline']['frac'],3), d['clocks'], d['check']['ok'], d['check']['fields_sha'])"; done
