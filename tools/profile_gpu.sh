set -x
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_c4_v2.csv python bench.py --workload c4 --steps 4 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_under_ncu_v2.log 2>&1
timeout 700 ncu --set full --clock-control none --import-source on -k regex:k_fused3d_tb2 -s 2 -c 1 -o gpurun_out/prof_tb2_r01_v2 -f python bench.py --workload c4 --steps 4 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_tb2_v2.log 2>&1
ls -la gpurun_out | tail -5
