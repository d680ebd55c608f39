timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py 2>&1 | tail -1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_n2_v11.json
python -c "
import json;d=json.load(open('gpurun_out/bench_n2_v11.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['phase_ms_per_step'],d['encode']['value'],d.get('train'),d['clocks'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 2>&1 | tail -1 | cut -c1-400
