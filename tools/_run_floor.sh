timeout 300 python -m pytest tests/test_loss_gpu.py -x -q 2>&1 | tail -5
timeout 600 python bench.py --steps 3 --warmup 3 --skip-cpu 2>&1 | tail -1 > gpurun_out/bench_loss.json
python -c "
import json;d=json.load(open('gpurun_out/bench_loss.json'));print(d['value'],d['ms_per_step'],d['loss'],d['encode']['ms_per_step'])"
