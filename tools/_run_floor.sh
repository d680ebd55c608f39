export PYTHONPATH=.
for b in 64 96 128 192 256 512; do timeout 120 python tools/encoder_probe.py bert $b 128 20 2>&1 | tail -1; done
timeout 600 python bench.py --steps 3 --warmup 3 --skip-cpu 2>&1 | tail -1 > gpurun_out/bench_train.json
python -c "
import json;d=json.load(open('gpurun_out/bench_train.json'));print(d['value'],d.get('train'),d['loss'])"
