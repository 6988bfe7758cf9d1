timeout 600 python -m pytest tests/test_gpu_bf16_accum.py tests/test_gpu_binned.py -q 2>&1 | tail -3
for b in 1 2 3 4 6 8; do timeout 300 python scripts/tune.py --only bwd --flags 0x20,0 --dtypes f32 --regimes R --batch $b; done
for b in 1 2 4; do timeout 300 python scripts/tune.py --only bwd --flags 0x20,0 --dtypes f32 --regimes R --batch $b --d-model 128; done
for b in 1 2 4 32; do timeout 300 python scripts/tune.py --only bwd --flags 0x20,0 --dtypes f32 --regimes R --batch $b --pyramid BBBC_520x696; done
