#!/bin/bash
# scratch: build library variants for on-GPU A/B of pipeline-synchronisation choices
set -e
cd "$(dirname "$0")/.."
build() { name=$1; shift; d=build/var_$name; mkdir -p $d
  for f in api search encoder loss; do nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -c openmatch_b200/csrc/$f.cu -o $d/$f.o & done; wait
  nvcc -shared -o $d/libom.so $d/*.o -gencode arch=compute_100a,code=sm_100a; echo built $d/libom.so; }
build v0
build v1 -DOM_SPIN_SLEEP_NS=32
build v2 -DOM_SPIN_SLEEP_NS=200
build v3 -DOM_SPIN_SLEEP_NS=32 -DOM_EPI_ALL_LANES_WAIT=500
build v4 -DOM_NO_PIPELINED_LD
build v5 -DOM_SPIN_SLEEP_NS=32 -DOM_EPI_POLL_SLEEP_NS=300
