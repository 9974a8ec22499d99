#!/bin/bash
# usage: scripts/gpu_ab2.sh "<variants under exp_lib/>" : A/B of experimental builds on config 3 / 5 (default two-pass path) and config 2
mkdir -p gpurun_out
run() { name=$1; lib=$2; shift 2
  export VX3_ENGINE_LIB=$PWD/exp_lib/libvx3_$lib.so
  timeout 300 python bench.py --warmup 2 --skip-cpu --skip-e2e "$@" 2>gpurun_out/err_$name.log | N=$name python -c "
import json,sys,os
d=json.loads(sys.stdin.read())
print(os.environ['N'], 'value %.3e'%d['value'], 'ms/step %.3f'%d['ms_per_step'], d['roofline']['kernel_ms'], flush=True)"
}
for v in $1; do
  run c3_$v $v --workload c3 --steps 10
  run c5_$v $v --workload c5 --sim-steps 100 --steps 3
done
