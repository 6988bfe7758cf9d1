timeout 900 python -m pytest tests/test_gpu_binned.py tests/test_gpu_fused.py -q -x 2>&1 | tail -8
timeout 300 python scripts/bench_fused.py 16
