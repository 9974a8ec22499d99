mkdir -p gpurun_out
export VX3_HALO_TIMEOUT_MS=2000
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 200 $T scripts/check_decomp_mp.py > gpurun_out/decomp_mp_ik.log 2>&1; tail -2 gpurun_out/decomp_mp_ik.log | cut -c1-300
for sf in 1 0; do
VX3_HALO_SENDFUSED=$sf timeout 150 $T bench.py --gpus 2 --workload c5 --steps 20 --warmup 3 --sim-steps 100 --skip-cpu --skip-e2e > gpurun_out/c5_n2_sf$sf.json 2> gpurun_out/c5_n2_sf$sf.err; tail -1 gpurun_out/c5_n2_sf$sf.err | cut -c1-300
python - <<PY
import json
try:
    l=json.loads(open("gpurun_out/c5_n2_sf$sf.json").read().strip().splitlines()[-1])
    print("sendfused=$sf", l["value"], l["ms_per_step"], l["selfcheck"]["bit_exact"], l["compute_us_per_sim_step"], l["halo_exposed_us_per_sim_step"], l["roofline"]["kernel_ms"])
except Exception as e:
    print("sendfused=$sf failed", e)
PY
done
python -m pytest tests/test_decomposition.py -m gpu -q -x 2>&1 | tail -3
