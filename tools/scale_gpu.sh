# usage (on the GPU box, through gpurun --gpus N):  bash tools/scale_gpu.sh N
# Strong scaling of the default workload on N GPUs: correctness first (bit-exact vs one engine, uneven slabs too),
# then the bench line with load-balanced and with equal slabs, and the per-rank sweep / period breakdown.
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
mkdir -p gpurun_out
timeout 300 $TR tests/multi_gpu_check.py --f32 2>&1 | grep -E "MULTI_GPU|rror|timed" | head -3
timeout 300 $TR tests/multi_gpu_check.py --f32 --balanced 2>&1 | grep -E "MULTI_GPU|rror|timed" | head -3
for off in 1 2; do timeout 300 $TR tests/multi_gpu_check.py --src-offset $off 2>&1 | grep -E "MULTI_GPU|rror|timed" | head -3; done
timeout 400 $TR bench.py --gpus $N --steps 100 --warmup 6 2>/dev/null | grep '^{' | tee gpurun_out/scale_n$N.json | cut -c1-300
FDTD_B200_BALANCE=0 timeout 400 $TR bench.py --gpus $N --steps 100 --warmup 6 --no-e2e 2>/dev/null | grep '^{' | tee gpurun_out/scale_n${N}_equal.json | cut -c1-300
FDTD_B200_SLAB_DEBUG=1 timeout 400 $TR bench.py --gpus $N --steps 100 --warmup 6 --no-e2e 2>&1 | grep "fdtd dbg" | sort | tail -$N
