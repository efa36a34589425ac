#!/usr/bin/env bash
# Round 2, 8-GPU diagnostics: per-rank sweep duration vs pair period (load-balanced and equal slabs).
cd "$(dirname "$0")/.." || exit 1
N=${1:-8}
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29651"
FDTD_B200_SLAB_DEBUG=1 timeout 400 $TR bench.py --gpus $N --steps 100 --warmup 6 --no-e2e --no-cpu --no-check 2>&1 | grep -E "fdtd dbg|^\{" | sed 's/"config".*"roofline"/.../' | cut -c1-220 | sort | tee $O/n${N}_slab_debug_balanced.log | tail -12
FDTD_B200_BALANCE=0 FDTD_B200_SLAB_DEBUG=1 timeout 400 $TR bench.py --gpus $N --steps 100 --warmup 6 --no-e2e --no-cpu --no-check 2>&1 | grep -E "fdtd dbg|^\{" | sed 's/"config".*"roofline"/.../' | cut -c1-220 | sort | tee $O/n${N}_slab_debug_equal.log | tail -12
timeout 400 $TR bench.py --gpus $N --steps 100 --warmup 6 --no-e2e --no-cpu --no-check --no-ops 2>&1 | grep -E "^\{" | cut -c1-160 | tee $O/n${N}_noops.log
