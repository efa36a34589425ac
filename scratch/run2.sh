timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
B="timeout 300 python bench.py --no-cpu --no-e2e --steps 40 --warmup 4"
pick() { python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); r=d['roofline']; print('RESULT', sys.argv[1], round(d['value']/1e9,2), 'frac', round(r['frac'],3), 'ms', round(d['ms_per_step'],4), 'kern', round(r.get('kernel_ms_per_step',0),4), 'post', round(r.get('post_ms_per_step',0),4))" "$1"; }
$B | pick c4_auto
FDTD_B200_TB2_ZONES=0 $B | pick c4_z0
FDTD_B200_FUSED_LX=256 $B | pick c4_lx256_z1
FDTD_B200_FUSED_LX=256 FDTD_B200_TB2_ZONES=0 $B | pick c4_lx256_z0
FDTD_B200_FUSED_LX=128 FDTD_B200_TB2_ZONES=0 $B | pick c4_lx128_z0
$B --workload 128x1024x1024 | pick slab128_auto
FDTD_B200_TB2_ZONES=1 $B --workload 128x1024x1024 | pick slab128_z1
$B --workload 128x1024x1024 --no-ops | pick slab128_noops
