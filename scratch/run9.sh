pick() { python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); r=d['roofline']; print('RESULT', sys.argv[1], round(d['value']/1e9,2), 'ms', round(d['ms_per_step'],4), d['clocks'])" "$1"; }
B="python bench.py --no-cpu --no-e2e --steps 60 --warmup 6 --workload 128x1024x1024 --no-ops"
CUDA_VISIBLE_DEVICES=1 timeout 300 $B | pick gpu1_alone
(CUDA_VISIBLE_DEVICES=0 timeout 300 $B | pick gpu0_both) &
(CUDA_VISIBLE_DEVICES=1 timeout 300 $B | pick gpu1_both) &
wait
