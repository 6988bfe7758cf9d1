timeout 200 python scripts/fused_profile.py 2>&1 | grep -E "msda|zero|Self CUDA time" | cut -c1-100,180-260 | tee gpurun_out/r02_fused_profile2.log
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_binned.py tests/test_encoder.py -x -q -m gpu 2>&1 | tail -3
