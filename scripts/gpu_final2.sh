#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --workload c3 --steps 10 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -c 200 gpurun_out/bench_c3.json; echo
timeout 400 python bench.py --workload c5 --steps 3 --warmup 1 --sim-steps 100 --skip-cpu > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; tail -c 200 gpurun_out/bench_c5.json; echo
timeout 300 python bench.py --workload c4 --steps 3 --warmup 1 --sim-steps 200 --skip-cpu > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; tail -c 200 gpurun_out/bench_c4.json; echo
timeout 400 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 300 gpurun_out/bench_c2.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 300 --csv --log-file gpurun_out/launches_c3.csv python bench.py --workload c3 --steps 1 --warmup 1 --sim-steps 100 --skip-cpu --skip-e2e > gpurun_out/ncu_c3.log 2>&1
VX3_LINK_QUEUE=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_links -s 40 -c 1 -o gpurun_out/prof_links_c3 -f python bench.py --workload c3 --steps 1 --warmup 1 --sim-steps 50 --skip-cpu --skip-e2e > gpurun_out/ncu_links.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_voxels -s 40 -c 1 -o gpurun_out/prof_voxels_c3 -f python bench.py --workload c3 --steps 1 --warmup 1 --sim-steps 50 --skip-cpu --skip-e2e > gpurun_out/ncu_voxels.log 2>&1
ls gpurun_out/*.ncu-rep
