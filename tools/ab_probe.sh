#!/bin/bash
# scratch: A/B the scan kernel variants on the GPU box
export PYTHONPATH=$PWD
echo "== A (pipelined LDTM, warp-wait)"; python tools/search_probe.py 1000000,6980,1000 8800000,6980,1000 2>&1 | grep -v property
echo "== B (no pipelined LDTM, warp-wait)"; OPENMATCH_B200_LIB=$PWD/build/variantB/libom_B.so python tools/search_probe.py 1000000,6980,1000 8800000,6980,1000 2>&1 | grep -v property
