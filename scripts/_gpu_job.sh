timeout 900 python -m pytest tests/test_gpu_binned.py -q -x 2>&1 | tail -2
timeout 300 python scripts/tune.py --only bwd --flags 0 --dtypes f32,bf16 --regimes R,U
timeout 300 python scripts/tune.py --only bwd --flags 0 --dtypes f32 --regimes R --d-model 128
timeout 300 python scripts/tune.py --only bwd --flags 0 --dtypes f32 --regimes R --pyramid BBBC_520x696 --batch 32
for b in 1 2 4; do timeout 300 python scripts/tune.py --only bwd --flags 0 --dtypes f32 --regimes R --batch $b; done
