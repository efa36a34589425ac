TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
pick() { python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); r=d['roofline']; print('RESULT', sys.argv[1], round(d['value']/1e9,2), 'frac', round(r['frac'],3), 'ms', round(d['ms_per_step'],4))" "$1"; }
B="bench.py --gpus 2 --no-e2e --steps 60 --warmup 6 --workload 256x1024x1024 --no-ops"
for k in 0 1 2 3 7; do FDTD_B200_DEBUG_SKIP=$k timeout 300 $TR $B 2>/dev/null | grep '^{' | pick skip$k; done
nvidia-smi topo -m | head -8
