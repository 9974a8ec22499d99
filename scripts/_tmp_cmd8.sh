mkdir -p gpurun_out
export VX3_HALO_TIMEOUT_MS=3000
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 200 $T scripts/check_decomp_mp.py > gpurun_out/decomp_mp_n8.log 2>&1; grep DECOMP_MP gpurun_out/decomp_mp_n8.log | cut -c1-200
run() {
  tag=$1; shift
  env "$@" timeout 200 $T bench.py --gpus 8 --workload c5 --steps 20 --warmup 3 --sim-steps 100 --skip-cpu --skip-e2e > gpurun_out/c5_n8_$tag.json 2> gpurun_out/c5_n8_$tag.err
  python - <<PY
import json
try:
    l=json.loads(open("gpurun_out/c5_n8_$tag.json").read().strip().splitlines()[-1])
    print("$tag", "%.4g" % l["value"], "%.2f us/step" % (10*l["ms_per_step"]), l["selfcheck"]["bit_exact"], "%.1f" % l["compute_us_per_sim_step"], "%.1f" % l["halo_exposed_us_per_sim_step"], l["roofline"]["kernel_ms"])
except Exception as e:
    print("$tag failed", e)
PY
}
run fused2 A=1
