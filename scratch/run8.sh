TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
timeout 300 $TR tests/multi_gpu_check.py 2>&1 | grep -E "MULTI_GPU|Error|error|timed" | head -5
timeout 300 $TR tests/multi_gpu_check.py --f32 2>&1 | grep -E "MULTI_GPU|Error|error|timed" | head -5
timeout 300 $TR tests/multi_gpu_check.py --sim 2>&1 | grep -E "MULTI_GPU|Error|error|timed" | head -5
B="bench.py --gpus 2 --no-e2e --steps 60 --warmup 6 --workload 256x1024x1024"
FDTD_B200_DEBUG_SKIP=8 timeout 300 $TR $B --no-ops 2>&1 | grep -E "fdtd dbg" | tail -2
FDTD_B200_DEBUG_SKIP=24 timeout 300 $TR $B --no-ops 2>&1 | grep -E "fdtd dbg" | tail -2
pick() { python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); r=d['roofline']; print('RESULT', sys.argv[1], round(d['value']/1e9,2), 'frac', round(r['frac'],3), 'ms', round(d['ms_per_step'],4))" "$1"; }
timeout 300 $TR $B --no-ops 2>/dev/null | grep '^{' | pick n2_256_noops
timeout 300 $TR $B 2>/dev/null | grep '^{' | pick n2_256
