#!/bin/bash
# First GPU call of the next round (one B200, ~1 min of run time): the two questions left open by round 1.
#   gpurun --timeout 600 -- 'bash tools/r02_first_call.sh'
# 1. scan filter variants (scan_epilogue.cuh, -DOM_SCAN_VARIANT=1|2|3): identical survivor sets? which shape is the
#    cheapest per survivor?  (variant 0 at 0 / 0.2 / 0.9 / 2.5 / 6.3 survivors per warp-tile: 1554 / 1477 / 1300 /
#    1053 / 880 TFLOP/s.)  Adopt the winner by building search.cu with -DOM_SCAN_VARIANT=<n>, then run
#    tests/test_search_gpu.py and bench.py.
# 2. the 2-CTA (cta_group::2) GEMM probe: bit-exact but ~770 TFLOP/s on every shape in round 1.
mkdir -p gpurun_out
timeout 200 build/selftest_gemm --scanvar > gpurun_out/r02_scanvar.log 2>&1; echo "rc=$?" >> gpurun_out/r02_scanvar.log
timeout 200 build/selftest_gemm --2sm > gpurun_out/r02_2sm.log 2>&1; echo "rc=$?" >> gpurun_out/r02_2sm.log
grep -h "scanvar\|perf\|rc=" gpurun_out/r02_scanvar.log | tail -40
grep -h "perf\|2sm:\|rc=" gpurun_out/r02_2sm.log | tail -20
