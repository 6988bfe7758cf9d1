timeout 600 python -m pytest tests/test_gpu_bf16_accum.py -q -s 2>&1 | grep -v "^$" | tail -30
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fused.py tests/test_encoder.py -m gpu -q -x 2>&1 | tail -4
timeout 300 python scripts/tune.py --only bwd --flags 0,0x80,0x8 --dtypes bf16 --regimes R,U
timeout 300 python scripts/tune.py --only bwd --flags 0,0x80 --dtypes bf16 --regimes R --d-model 128
timeout 300 python scripts/tune.py --only bwd --flags 0,0x80 --dtypes bf16 --regimes R --pyramid BBBC_520x696 --batch 32
