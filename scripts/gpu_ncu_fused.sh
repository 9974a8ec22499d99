#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused -s 10 -c 1 -o gpurun_out/prof_fused_c5 -f python bench.py --workload c5 --steps 1 --warmup 0 --sim-steps 16 --skip-cpu --skip-e2e > gpurun_out/ncu_fused_c5.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused -s 40 -c 1 -o gpurun_out/prof_fused_c3 -f python bench.py --workload c3 --steps 1 --warmup 1 --sim-steps 50 --skip-cpu --skip-e2e > gpurun_out/ncu_fused_c3.log 2>&1
ls -la gpurun_out/*.ncu-rep
