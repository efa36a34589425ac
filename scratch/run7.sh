TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
B="bench.py --gpus 2 --no-e2e --steps 60 --warmup 6 --workload 256x1024x1024 --no-ops"
FDTD_B200_DEBUG_SKIP=8 timeout 300 $TR $B 2>&1 | grep -E "fdtd dbg|ms_per_step" | cut -c1-200
FDTD_B200_DEBUG_SKIP=15 timeout 300 $TR $B 2>&1 | grep -E "fdtd dbg" | cut -c1-200
python bench.py --no-cpu --no-e2e --steps 60 --warmup 6 --workload 128x1024x1024 --no-ops | cut -c1-400
