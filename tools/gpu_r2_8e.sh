#!/usr/bin/env bash
# Round 2, 8-GPU: the named config 5 (2048x1024x512, coupler + Drude pad + uniaxial cladding + 2-port S-parameters) with
# device-painted coefficients on 8 x-slabs; per-component-Cb slabs bit-exact against one engine on real separate GPUs.
cd "$(dirname "$0")/.." || exit 1
N=${1:-8}
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29671"
timeout 200 $TR tests/multi_gpu_check.py --aniso 2>&1 | grep -E "MULTI_GPU|rror|MISMATCH" | head -3 | tee $O/n${N}e_multi_gpu_check_aniso.log
timeout 400 $TR bench.py --gpus $N --workload c5 --steps 20 --warmup 3 --no-cpu --no-e2e 2> $O/n${N}e_bench_c5_aniso.err | grep '^{' > $O/n${N}e_bench_c5_aniso.json
tail -1 $O/n${N}e_bench_c5_aniso.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['value']/1e9,2),'Gcell/s', 'ms/step', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],3), d['clocks'], d['check']['ok'], d['check']['fields_sha'], d['check'].get('timed_fields_sha'), d.get('setup'), d.get('s_params',{}).get('S21'))"
tail -3 $O/n${N}e_bench_c5_aniso.err
