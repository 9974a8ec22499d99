#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29541 bench.py --gpus 8 --workload c3 --steps 10 --warmup 3 --skip-cpu > gpurun_out/bench_c3_n8.json 2> gpurun_out/bench_c3_n8.err
timeout 400 $TR --master-port 29542 bench.py --gpus 8 --workload c5 --sim-steps 100 --steps 3 --warmup 1 --skip-cpu > gpurun_out/bench_c5_n8.json 2> gpurun_out/bench_c5_n8.err
timeout 200 $TR --master-port 29543 bench.py --gpus 8 --steps 20 --warmup 3 --skip-cpu > gpurun_out/bench_c2_n8.json 2> gpurun_out/bench_c2_n8.err
wc -l gpurun_out/bench_c3_n8.json gpurun_out/bench_c5_n8.json gpurun_out/bench_c2_n8.json
