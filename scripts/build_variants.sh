#!/bin/bash
# developer A/B helper: build experimental engine variants into exp_lib/ (git-ignored; travels with gpurun)
# usage: scripts/build_variants.sh name1:"-DX=1 -DY=2" name2:"..."
set -e
cd "$(dirname "$0")/.."
mkdir -p exp_lib
P=voxcraft-sim_b200/csrc
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  nvcc -std=c++17 -O3 -lineinfo -shared -Xcompiler -fPIC,-O2,-ffp-contract=off -I include -I $P -gencode arch=compute_100a,code=sm_100a \
    -fmad=false $flags -o exp_lib/libvx3_$name.so $P/engine/vx3_engine.cu $P/host/vx3_materials.cpp $P/host/vx3_builder.cpp $P/host/vx3_xml.cpp $P/host/vx3_vxa.cpp $P/host/vx3_worker.cpp &
done
wait
ls -la exp_lib
