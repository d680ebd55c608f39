timeout 600 python -m pytest tests/test_search_gpu.py -x -q 2>&1 | tail -5
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v10.csv python bench.py --steps 1 --warmup 1 --skip-cpu --skip-encode > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:loss_fused -c 2 -o gpurun_out/loss_full_v10 python tools/loss_probe.py > gpurun_out/ncu_loss.log 2>&1
tail -2 gpurun_out/ncu_loss.log
