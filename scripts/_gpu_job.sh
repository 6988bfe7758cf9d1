timeout 900 python -m pytest tests/test_gpu_binned.py -q -x 2>&1 | tail -3
timeout 300 python scripts/tune.py --only bwd --flags 0x20,0x40 --dtypes f32 --regimes R,U
for v in r_noloop r_nosweep; do
  echo "== variant $v"
  PCTRANS_MSDA_LIB=pctrans_b200/_build/variants/libpctrans_msda_$v.so timeout 300 python scripts/tune.py --only bwd --flags 0x40 --dtypes f32 --regimes R
done
timeout 300 python scripts/tune.py --only bwd --flags 0x20,0x40 --dtypes f32,bf16 --regimes R --d-model 128
timeout 300 python scripts/tune.py --only bwd --flags 0x20,0x40 --dtypes bf16 --regimes R
timeout 300 python scripts/tune.py --only bwd --flags 0x20,0x40 --dtypes f32 --regimes R --pyramid BBBC_520x696 --batch 32
timeout 300 python scripts/tune.py --only bwd --flags 0x20,0x40 --dtypes f32 --regimes R --d-model 512
