#!/bin/bash
# A/B of compile-time variants of dp_temporal.cu ON the GPU box (like ab_frame.sh): rebuilds the object with each flag set, relinks
# libdp_engine.so and times scripts/one_frame.py.  Usage: scripts/ab_temporal.sh "<flags A>" "<flags B>" ...  (leaves the LAST variant built)
cd "$(dirname "$0")/.."
CS=dragposer_b200/csrc; OBJ=dragposer_b200/build
for flags in "$@"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden $flags -c -o $OBJ/dp_temporal.o $CS/dp_temporal.cu || exit 1
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o dragposer_b200/libdp_engine.so $OBJ/*.o || exit 1
  for r in 1 2 3; do echo "[$flags] $(python scripts/one_frame.py 4096 12 | tail -1)"; done
done
