#!/usr/bin/env bash
# Round 2, 8-GPU: scaling line after the early halo push + per-rank breakdown.
cd "$(dirname "$0")/.." || exit 1
N=${1:-8}
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29671"
timeout 300 $TR tests/multi_gpu_check.py --f32 2>&1 | grep -E "MULTI_GPU|rror|MISMATCH" | head -2
FDTD_B200_SLAB_DEBUG=1 timeout 400 $TR bench.py --gpus $N --steps 100 --warmup 6 --no-e2e --no-cpu --no-check 2>&1 | grep -E "fdtd dbg.*pairs 50|^\{" | cut -c1-170 | sort | tee $O/n${N}d_slab_debug.log
timeout 500 $TR bench.py --gpus $N --steps 100 --warmup 6 --no-cpu 2> $O/n${N}d_bench_c4.err | grep '^{' > $O/n${N}d_bench_c4.json
tail -1 $O/n${N}d_bench_c4.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['value']/1e9,2),'Gcell/s', 'ms/step', round(d['ms_per_step'],4), 'e2e', d['e2e'] and round(d['e2e']['value']/1e9,2), d['clocks'], d['check']['ok'], d['check']['fields_sha'])"
