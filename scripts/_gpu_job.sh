timeout 900 python -m pytest tests/test_gpu_binned.py tests/test_gpu_fused.py tests/test_gpu_parity.py tests/test_encoder.py -q -x 2>&1 | tail -3
timeout 300 python scripts/tune.py --only bwd --flags 0 --dtypes f32 --regimes R
timeout 300 python scripts/bench_encoder.py --mode train --batch 16 2>&1 | tail -1 | cut -c1-330
