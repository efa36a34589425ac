#!/usr/bin/env bash
# Round 2, GPU session J (2 GPUs): heterogeneous slabs + slab ADE, bit-exact vs one engine; self-checking multi-GPU bench.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
( time timeout 1200 python -m pytest tests/test_multi_slab.py tests/test_gpu_ade_fused.py -m gpu -q --timeout 900 ) > $O/j_pytest_multi.log 2>&1; tail -12 $O/j_pytest_multi.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631"
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu > $O/j_bench_n2.json 2> $O/j_bench_n2.err; tail -c 1500 $O/j_bench_n2.json; tail -3 $O/j_bench_n2.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu --no-e2e > $O/j_bench_n1.json 2> $O/j_bench_n1.err; tail -c 700 $O/j_bench_n1.json
