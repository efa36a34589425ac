#!/usr/bin/env bash
# Round 2, GPU session Y: ncu --set full of the index-coded heterogeneous sweep (what bounds it now that the coefficient
# streams are gone).
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 100 ncu --set full --clock-control none --import-source on -k regex:"k_fused3d_het" -c 3 -f -o $O/y_het_indexed python bench.py --workload c3 --aniso --steps 4 --warmup 3 --no-cpu --no-e2e --no-check > $O/y_ncu_full.log 2>&1
ls -la $O/y_het_indexed.ncu-rep; tail -2 $O/y_ncu_full.log | cut -c1-200
