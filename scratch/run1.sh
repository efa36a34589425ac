set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
B="timeout 300 python bench.py --no-cpu --no-e2e --steps 40 --warmup 4"
pick() { python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('RESULT', sys.argv[1], round(d['value']/1e9,2), 'frac', round(d['roofline']['frac'],3), 'ms', round(d['ms_per_step'],4))" "$1"; }
$B | pick auto
for lx in 512 342 256 128 64; do FDTD_B200_FUSED_LX=$lx $B | pick lx$lx; done
$B --no-ops | pick noops
$B --workload 128x1024x1024 | pick slab128
$B --workload 128x1024x1024 --no-ops | pick slab128_noops
FDTD_B200_FUSED_LX=43 $B --workload 128x1024x1024 | pick slab128_lx43
FDTD_B200_FUSED_LX=128 $B --workload 128x1024x1024 | pick slab128_lx128
