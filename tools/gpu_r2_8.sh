#!/usr/bin/env bash
# Round 2, 8-GPU session: bit-exactness on 8 devices (uniform + heterogeneous slabs), c4 scaling line with self-check, c5 line.
cd "$(dirname "$0")/.." || exit 1
N=${1:-8}
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29641"
for a in "--f32" "--het" "--het --f32" "--sim"; do timeout 300 $TR tests/multi_gpu_check.py $a 2>&1 | grep -E "MULTI_GPU|rror|timed|MISMATCH" | head -3; done | tee $O/n${N}_multi_gpu_check.log
timeout 500 $TR bench.py --gpus $N --steps 100 --warmup 6 --no-cpu 2> $O/n${N}_bench_c4.err | grep '^{' > $O/n${N}_bench_c4.json
timeout 700 $TR bench.py --gpus $N --workload c5 --steps 40 --warmup 4 --no-cpu --no-e2e 2> $O/n${N}_bench_c5.err | grep '^{' > $O/n${N}_bench_c5.json
for f in $O/n${N}_bench_c4.json $O/n${N}_bench_c5.json; do echo "== $f"; tail -1 $f | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print(round(d['value']/1e9,2),'Gcell/s frac',round(d['roofline']['frac'],3), 'ms/step', round(d['ms_per_step'],3), 'e2e', d['e2e'] and round(d['e2e']['value']/1e9,2), d['clocks'], d.get('check'), d.get('s_params') and d['s_params']['S21'])
except Exception as e: print('unparsed', e)"; tail -2 ${f%.json}.err; done
