#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29551 bench.py --gpus 8 --workload c3 --steps 10 --warmup 3 --skip-cpu --skip-e2e > gpurun_out/bench_c3_n8b.json 2> gpurun_out/bench_c3_n8b.err
python -c "
import json
d=json.load(open('gpurun_out/bench_c3_n8b.json')); print('c3 n8 value %.3e ms/step %.3f'%(d['value'],d['ms_per_step']), d['clocks'])"
