timeout 300 build/selftest_gemm --2sm > gpurun_out/r02_2sm_product.log 2>&1; echo rc=$? >> gpurun_out/r02_2sm_product.log
grep -v "^   " gpurun_out/r02_2sm_product.log | tail -32
timeout 900 python -m pytest tests/test_search_gpu.py -q -x > gpurun_out/r02_tests_pair.log 2>&1; tail -5 gpurun_out/r02_tests_pair.log
OM_PROFILE=1 OM_PAIR=1 python tools/search_probe.py 8800000,6980,1000 > gpurun_out/r02_search_probe_pair1.log 2>&1; tail -6 gpurun_out/r02_search_probe_pair1.log
OM_PROFILE=1 OM_PAIR=0 python tools/search_probe.py 8800000,6980,1000 > gpurun_out/r02_search_probe_pair0.log 2>&1; tail -6 gpurun_out/r02_search_probe_pair0.log
