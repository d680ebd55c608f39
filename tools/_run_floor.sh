ulimit -c 0
timeout 150 build/selftest_gemm --2sm > gpurun_out/selftest_2sm.log 2>&1; echo "rc=$?" >> gpurun_out/selftest_2sm.log
tail -25 gpurun_out/selftest_2sm.log
