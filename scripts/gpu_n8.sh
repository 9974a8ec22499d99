#!/bin/bash
# 8-GPU lines: config 3 (weak: 512 robots per GPU), config 5 (strong: one body in 8 slabs), default bench (config 2 replicas)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 bench.py --gpus 8 --workload c3 --steps 10 --warmup 3 --skip-cpu > gpurun_out/bench_c3_n8.json 2> gpurun_out/bench_c3_n8.err; tail -c 900 gpurun_out/bench_c3_n8.json; echo
timeout 900 $TR --master-port 29512 bench.py --gpus 8 --workload c5 --sim-steps 100 --steps 3 --warmup 1 --skip-cpu > gpurun_out/bench_c5_n8.json 2> gpurun_out/bench_c5_n8.err; tail -c 900 gpurun_out/bench_c5_n8.json; echo
timeout 600 $TR --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_c2_n8.json 2> gpurun_out/bench_c2_n8.err; tail -c 600 gpurun_out/bench_c2_n8.json; echo
tail -3 gpurun_out/bench_c5_n8.err
