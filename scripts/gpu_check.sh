#!/bin/bash
# one gpurun call: full GPU suite, e2e breakdown, default benches
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log
python scripts/e2e_breakdown.py c3 2>&1 | tail -16
python scripts/e2e_breakdown.py c2 2>&1 | tail -14
