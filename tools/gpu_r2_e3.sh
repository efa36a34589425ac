#!/usr/bin/env bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
for d in 0 1 2 4 3 5 6 7; do
  echo "== dbg $d"; FDTD_B200_YEEX_DBG=$d timeout 120 python - <<'PY' 2>&1 | tail -2
import numpy as np, sys
sys.path.insert(0,'.')
import prismo_b200 as pb
from prismo_b200 import _lib
d=2e-8; dt=0.5*d/(299792458.0*np.sqrt(3))
eng=pb.Engine(3,(21,29,70),(d,)*3,dt,dtype="float32",flags=_lib.FLAG_YEE)
rng=np.random.default_rng(0)
for c in ("Ex","Ey","Ez","Hx","Hy","Hz"): eng.upload(c, rng.standard_normal(eng.field_shape(c)))
try:
    eng.run(2); eng.sync(); print("ran ok")
except Exception as ex: print("FAIL", str(ex)[-60:])
PY
done
