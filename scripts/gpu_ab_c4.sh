#!/bin/bash
mkdir -p gpurun_out
for v in $1; do
  export VX3_ENGINE_LIB=$PWD/exp_lib/libvx3_$v.so
  python bench.py --warmup 2 --skip-cpu --skip-e2e --workload c4 --sim-steps 200 --steps 3 2>gpurun_out/err.log | V=$v python -c "
import json,sys,os
d=json.loads(sys.stdin.read())
print(os.environ['V'], 'value %.3e'%d['value'], 'ms/step %.3f'%d['ms_per_step'], d['roofline']['kernel_ms'])"
done
