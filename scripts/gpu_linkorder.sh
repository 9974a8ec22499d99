#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
bash scripts/ab_env.sh VX3_LINK_ORDER "0 1" "c3 --steps 10" "c5 --sim-steps 100 --steps 3" "c4 --sim-steps 200 --steps 3" "c2 --no-persistent --steps 5"
