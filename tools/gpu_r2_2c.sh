#!/usr/bin/env bash
# Round 2, 2 real GPUs: config 5 with the index-coded heterogeneous sweep on x-slabs (digest must equal the 1-GPU lines),
# then the slab check with index-coded slabs against an array-path whole engine.
cd "$(dirname "$0")/.." || exit 1
N=${1:-2}
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29671"
timeout 100 $TR bench.py --gpus $N --workload c5 --steps 20 --warmup 3 --no-cpu --no-e2e 2> $O/n${N}c_bench_c5_indexed.err | grep '^{' > $O/n${N}c_bench_c5_indexed.json
tail -1 $O/n${N}c_bench_c5_indexed.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['value']/1e9,2),'Gcell/s', 'ms/step', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],3), d['clocks'], d['check']['ok'], d['check']['fields_sha'], d['check'].get('timed_fields_sha'), d.get('setup'))"
tail -2 $O/n${N}c_bench_c5_indexed.err
timeout 60 $TR tests/multi_gpu_check.py --aniso --indexed 2>&1 | grep -E "MULTI_GPU|rror|MISMATCH" | head -3 | tee $O/n${N}c_multi_gpu_check_indexed.log
