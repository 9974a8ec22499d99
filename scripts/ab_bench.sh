#!/bin/bash
# developer A/B helper: run bench.py for each experimental engine build under exp_lib/ (VX3_ENGINE_LIB override)
# usage: scripts/ab_bench.sh "<lib suffixes>" "<workload args>" ...
libs="$1"; shift
for v in $libs; do
  if [ "$v" = "product" ]; then unset VX3_ENGINE_LIB; else export VX3_ENGINE_LIB=$PWD/exp_lib/libvx3_$v.so; fi
  for w in "$@"; do
    python bench.py --warmup 1 --skip-cpu --skip-e2e --workload $w 2>gpurun_out/err.log | V=$v python -c "
import json,sys,os
d=json.loads(sys.stdin.read())
print(os.environ['V'], d['config']['workload'][:8], 'value %.3e'%d['value'], 'ms/step %.3f'%d['ms_per_step'], d['roofline']['kernel_ms'])"
    grep -i "persist timing" gpurun_out/err.log | tail -1
  done
done
