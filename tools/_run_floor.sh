timeout 900 python -m pytest tests/test_encoder_gpu.py tests/test_search_gpu.py -q 2>&1 | tail -12
export PYTHONPATH=.
timeout 120 python tools/encoder_probe.py bert 64 512 10 2>&1 | tail -1
timeout 120 python tools/encoder_probe.py bert 128 256 10 2>&1 | tail -1
timeout 120 python tools/encoder_probe.py bert 256 128 10 2>&1 | tail -1
