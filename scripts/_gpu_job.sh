timeout 900 python -m pytest tests/test_gpu_linear.py tests/test_encoder.py -q -x 2>&1 | tail -4
timeout 300 python scripts/bench_encoder.py --mode train --batch 16 2>&1 | tail -1 | cut -c1-420
timeout 300 python scripts/bench_encoder.py --mode train --pyramid bbbc --batch 32 2>&1 | tail -1 | cut -c1-420
timeout 300 python scripts/bench_encoder.py --mode train --batch 16 --d-model 128 2>&1 | tail -1 | cut -c1-420
timeout 300 python scripts/profile_layer.py 2>&1 | grep -E "layernorm|layer_norm|GammaBeta|Self CUDA time"
