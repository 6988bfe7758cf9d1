timeout 900 python -m pytest tests/test_gpu_binned.py tests/test_gpu_fused.py tests/test_gpu_bounds.py -q -x 2>&1 | tail -3
timeout 300 python scripts/tune.py --only bwd --flags 0 --dtypes f32,bf16 --regimes R,U
timeout 300 python scripts/tune.py --only bwd --flags 0 --dtypes f32,bf16 --regimes R --d-model 128
timeout 300 python scripts/tune.py --only bwd --flags 0 --dtypes f32 --regimes R --d-model 512
timeout 300 python scripts/tune.py --only bwd --flags 0 --dtypes f32 --regimes R --pyramid BBBC_520x696 --batch 32
