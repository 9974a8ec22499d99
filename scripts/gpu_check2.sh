#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
for w in "c3 --steps 10" "c5 --sim-steps 100 --steps 3" "c4 --sim-steps 200 --steps 3" "c2 --no-persistent --steps 5"; do
  python bench.py --warmup 2 --skip-cpu --skip-e2e --workload $w 2>gpurun_out/err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['config']['workload'][:8], 'value %.3e'%d['value'], 'ms/step %.3f'%d['ms_per_step'], d['roofline']['kernel_ms'])"
done
