#!/bin/bash
export PYTHONPATH=$PWD
for v in v0 v1 v2 v3 v4 v5; do echo "== $v"; OPENMATCH_B200_LIB=$PWD/build/var_$v/libom.so OM_GROWTH=${OM_GROWTH:-2} OM_ITERS=5 python tools/search_probe.py 8800000,6980,1000 2>&1 | grep -v property; done
