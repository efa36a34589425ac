TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29611"
timeout 120 $TR tests/multi_gpu_check.py --f32 2>&1 | grep -E "MULTI_GPU|Error|error|timed" | head -3
timeout 200 $TR bench.py --gpus 4 --no-e2e --steps 100 --warmup 6 2>/dev/null | grep '^{' > gpurun_out/scale2_n4.json; cut -c1-330 gpurun_out/scale2_n4.json
