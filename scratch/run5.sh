TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
timeout 300 $TR tests/multi_gpu_check.py --f32 2>&1 | grep -E "MULTI_GPU|Error|error|timed" | head -5
pick() { python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); r=d['roofline']; print('RESULT', sys.argv[1], round(d['value']/1e9,2), 'frac', round(r['frac'],3), 'ms', round(d['ms_per_step'],4))" "$1"; }
B="bench.py --gpus 2 --no-e2e --steps 60 --warmup 6"
timeout 300 $TR $B --workload 256x1024x1024 2>/dev/null | grep '^{' | pick n2_256
timeout 300 $TR $B --workload 256x1024x1024 --no-ops 2>/dev/null | grep '^{' | pick n2_256_noops
FDTD_B200_TB2=0 timeout 300 $TR $B --workload 256x1024x1024 2>/dev/null | grep '^{' | pick n2_256_onestep
