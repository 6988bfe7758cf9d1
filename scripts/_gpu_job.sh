timeout 900 python -m pytest tests/test_gpu_linear.py tests/test_gpu_fused.py tests/test_reference_module_parity.py -q -x 2>&1 | tail -6
