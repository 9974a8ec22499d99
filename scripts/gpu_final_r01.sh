#!/bin/bash
# final round-1 evidence: GPU suite, bench lines (default + reference arm + other configs), ncu launch lists
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log; grep "full run" gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 400 gpurun_out/bench_c2.json; echo
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_c2_ref.json 2> gpurun_out/bench_c2_ref.err; tail -c 300 gpurun_out/bench_c2_ref.json; echo
timeout 300 python bench.py --workload c3 --steps 10 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -c 300 gpurun_out/bench_c3.json; echo
timeout 400 python bench.py --workload c5 --steps 3 --warmup 1 --sim-steps 100 --skip-cpu > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; tail -c 300 gpurun_out/bench_c5.json; echo
timeout 300 python bench.py --workload c4 --steps 3 --warmup 1 --sim-steps 200 --skip-cpu > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; tail -c 300 gpurun_out/bench_c4.json; echo
timeout 300 python bench.py --workload c3 --steps 10 --warmup 3 --fused --skip-cpu > gpurun_out/bench_c3_fused.json 2> gpurun_out/bench_c3_fused.err
timeout 400 python bench.py --workload c5 --steps 3 --warmup 1 --sim-steps 100 --skip-cpu --fused > gpurun_out/bench_c5_fused.json 2> gpurun_out/bench_c5_fused.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 2 --warmup 1 --skip-cpu --skip-e2e > gpurun_out/ncu_c2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 300 --csv --log-file gpurun_out/launches_c3.csv python bench.py --workload c3 --steps 1 --warmup 1 --sim-steps 100 --skip-cpu --skip-e2e > gpurun_out/ncu_c3.log 2>&1
ls gpurun_out | head -50
