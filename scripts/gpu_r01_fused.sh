#!/bin/bash
# one gpurun call: fused-step tests first, then A/B benches fused vs two-pass, then the rest of the GPU suite
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused.py -x -q -s > gpurun_out/pytest_fused.log 2>&1; tail -15 gpurun_out/pytest_fused.log
for w in c3 c5; do
  extra=""; [ $w = c5 ] && extra="--sim-steps 100 --steps 3"; [ $w = c3 ] && extra="--steps 10"
  timeout 300 python bench.py --workload $w --warmup 3 --skip-cpu --skip-e2e $extra > gpurun_out/bench_${w}_fused.json 2> gpurun_out/bench_${w}_fused.err
  timeout 300 python bench.py --workload $w --warmup 3 --skip-cpu --skip-e2e --no-fused $extra > gpurun_out/bench_${w}_twopass.json 2> gpurun_out/bench_${w}_twopass.err
  for v in fused twopass; do python - <<PY
import json
d=json.load(open("gpurun_out/bench_${w}_$v.json"))
print("$w $v value %.3e ms/step %.3f"%(d["value"],d["ms_per_step"]), d["roofline"]["kernel"], "frac %.3f"%d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["config"].get("fused_step"))
PY
  done
done
VX3_CREATE_TIMING=1 timeout 300 python bench.py --workload c5 --warmup 0 --steps 1 --sim-steps 10 --skip-cpu --skip-e2e 2>&1 >/dev/null | grep "create timing" | head -20
timeout 300 python bench.py --workload c2 --no-persistent --steps 5 --skip-cpu --skip-e2e > gpurun_out/bench_c2_stream_fused.json 2>gpurun_out/err.log; python -c "
import json; d=json.load(open('gpurun_out/bench_c2_stream_fused.json')); print('c2 streaming(fused) ms/step %.3f'%d['ms_per_step'], d['roofline']['kernel_ms'])"
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
