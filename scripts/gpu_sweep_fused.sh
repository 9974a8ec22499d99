#!/bin/bash
# A/B: fused-step tile parameters and the link pass at 5 CTAs/SM (experimental builds under exp_lib/)
mkdir -p gpurun_out
run() { # name lib extra-args...
  name=$1; lib=$2; shift 2
  if [ "$lib" = product ]; then unset VX3_ENGINE_LIB; else export VX3_ENGINE_LIB=$PWD/exp_lib/libvx3_$lib.so; fi
  timeout 300 python bench.py --warmup 2 --skip-cpu --skip-e2e "$@" 2>gpurun_out/err_$name.log | N=$name python -c "
import json,sys,os
d=json.loads(sys.stdin.read())
print(os.environ['N'], 'value %.3e'%d['value'], 'ms/step %.3f'%d['ms_per_step'], d['roofline']['kernel_ms'], flush=True)"
}
for v in base t96 bv248 bv88c5 bv60; do
  run c3_$v $v --workload c3 --steps 10
  run c5_$v $v --workload c5 --sim-steps 100 --steps 3
done
for v in base l5; do
  run c3_twopass_$v $v --workload c3 --steps 10 --no-fused
  run c5_twopass_$v $v --workload c5 --sim-steps 100 --steps 3 --no-fused
done
