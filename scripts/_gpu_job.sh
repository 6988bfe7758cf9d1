timeout 900 python -m pytest tests/test_gpu_binned.py -q -x 2>&1 | tail -5
