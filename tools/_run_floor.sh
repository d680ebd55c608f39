timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py 2>&1 | tail -1 > gpurun_out/bench_n1_v11.json
python -c "
import json;d=json.load(open('gpurun_out/bench_n1_v11.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['frac'],d['encode']['value'],d['encode'].get('t5_base_gtr'),d.get('train',{}).get('ms_per_step'),d['clocks'],d['gpu_launches'],d['cpu_baseline']['value'])"
OM_D=1024 OM_PROFILE=1 timeout 300 python tools/search_probe.py 2625000,6980,1000 2>&1 | tail -4
