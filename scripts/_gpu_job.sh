mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fused.py -q -x 2>&1 | tail -4
echo "== smem taps (release)"
timeout 300 python scripts/tune.py --only fwd --flags 0 --dtypes f32,bf16 --regimes R,U
timeout 300 python scripts/tune.py --only fwd --flags 0 --dtypes f32,bf16 --regimes R --d-model 128
echo "== shuffles (variant)"
export PCTRANS_MSDA_LIB=pctrans_b200/_build/variants/libpctrans_msda_fwd_shfl.so
timeout 300 python scripts/tune.py --only fwd --flags 0 --dtypes f32,bf16 --regimes R,U
timeout 300 python scripts/tune.py --only fwd --flags 0 --dtypes f32,bf16 --regimes R --d-model 128
