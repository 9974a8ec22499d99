#!/bin/bash
# one gpurun call: GPU tests, benches, ncu launch lists and full captures of the dominant kernels (summaries -> profiles/)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
VX3_PERSIST_TIMING=1 timeout 300 python bench.py --steps 5 --warmup 1 --skip-cpu --skip-e2e > gpurun_out/bench_c2_timing.json 2> gpurun_out/bench_c2_timing.err
grep "persist timing" gpurun_out/bench_c2_timing.err | tail -3
timeout 300 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; cat gpurun_out/bench_c2.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_c2_ref.json 2> gpurun_out/bench_c2_ref.err; cat gpurun_out/bench_c2_ref.json
timeout 300 python bench.py --workload c3 --steps 10 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; cat gpurun_out/bench_c3.json
timeout 300 python bench.py --workload c5 --steps 3 --warmup 1 --sim-steps 100 --skip-cpu > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; cat gpurun_out/bench_c5.json
timeout 300 python bench.py --workload c4 --steps 3 --warmup 1 --sim-steps 200 --skip-cpu > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; cat gpurun_out/bench_c4.json; tail -2 gpurun_out/bench_c4.err
# launch lists (same command as the bench, fewer steps)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 2 --warmup 1 --skip-cpu --skip-e2e > gpurun_out/ncu_c2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 300 --csv --log-file gpurun_out/launches_c3.csv python bench.py --workload c3 --steps 1 --warmup 1 --sim-steps 100 --skip-cpu --skip-e2e > gpurun_out/ncu_c3.log 2>&1
# full captures
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_links -s 40 -c 2 -o gpurun_out/prof_links_c3 -f python bench.py --workload c3 --steps 1 --warmup 1 --sim-steps 50 --skip-cpu --skip-e2e > gpurun_out/ncu_links.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_voxels -s 40 -c 2 -o gpurun_out/prof_voxels_c3 -f python bench.py --workload c3 --steps 1 --warmup 1 --sim-steps 50 --skip-cpu --skip-e2e > gpurun_out/ncu_voxels.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_persistent -s 1 -c 1 -o gpurun_out/prof_persistent_c2 -f python bench.py --steps 1 --warmup 1 --sim-steps 200 --skip-cpu --skip-e2e > gpurun_out/ncu_persist.log 2>&1
VX3_LINK_QUEUE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_links_deferred -s 4 -c 1 -o gpurun_out/prof_links_deferred_c5 -f python bench.py --workload c5 --steps 1 --warmup 0 --sim-steps 8 --skip-cpu --skip-e2e > gpurun_out/ncu_links_c5.log 2>&1
ls -la gpurun_out
