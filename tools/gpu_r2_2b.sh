#!/usr/bin/env bash
# Round 2, 2-GPU session: slab tests after the early-push change (ghost sources at offsets 0, 1, 2 keep the late push).
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
( time timeout 1200 python -m pytest tests/test_multi_slab.py -m gpu -q --timeout 900 ) > $O/t2_pytest_multi.log 2>&1; echo "PYTEST: $(grep -E ' passed| failed' $O/t2_pytest_multi.log | tail -1)"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29661"
for off in 0 1 2 3 4; do timeout 300 $TR tests/multi_gpu_check.py --src-offset $off 2>&1 | grep -E "MULTI_GPU|rror|MISMATCH" | head -2; done
FDTD_B200_SLAB_DEBUG=1 timeout 400 $TR bench.py --gpus 2 --steps 40 --warmup 4 --no-e2e --no-cpu 2>&1 | grep -E "fdtd dbg.*pairs 2[0-9]|^\{" | cut -c1-200
