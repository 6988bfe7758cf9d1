mkdir -p gpurun_out
export PCTRANS_MSDA_LIB=pctrans_b200/_build/variants/libpctrans_msda_t1024pf0.so
timeout 300 python scripts/tune.py --only bwd --flags 0x20000,0x20 --dtypes f32,bf16 --regimes R --d-model 128
timeout 300 python scripts/tune.py --only bwd --flags 0x20000,0x20 --dtypes bf16 --regimes R
timeout 300 python scripts/tune.py --only bwd --flags 0x20000,0x20 --dtypes f32 --regimes U
timeout 300 python scripts/tune.py --only bwd --flags 0x20000,0x20 --dtypes f32 --regimes R --d-model 512
timeout 300 python scripts/tune.py --only bwd --flags 0x20000,0x20 --dtypes f32 --regimes R --pyramid BBBC_520x696 --batch 32
timeout 300 python scripts/tune.py --only bwd --flags 0x20000,0x20 --dtypes f32 --regimes R --batch 2
