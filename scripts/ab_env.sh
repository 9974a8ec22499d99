#!/bin/bash
# developer A/B helper: bench.py under different values of one environment variable
# usage: scripts/ab_env.sh VAR "v1 v2 ..." "<workload args>" ...
var="$1"; vals="$2"; shift 2
for v in $vals; do
  for w in "$@"; do
    env $var=$v python bench.py --warmup 1 --skip-cpu --skip-e2e --workload $w 2>gpurun_out/err.log | V="$var=$v" python -c "
import json,sys,os
d=json.loads(sys.stdin.read())
print(os.environ['V'], d['config']['workload'][:8], 'value %.3e'%d['value'], 'ms/step %.3f'%d['ms_per_step'], d['roofline']['kernel_ms'])"
  done
done
