mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_binned.py -q 2>&1 | tail -15
echo "== release 768"
timeout 300 python scripts/tune.py --only bwd --flags 0,0x20000,0x40000,0x20 --dtypes f32 --regimes R
for v in t1024 t896 t640 c2t512 c2t384; do
  echo "== variant $v"
  PCTRANS_MSDA_LIB=pctrans_b200/_build/variants/libpctrans_msda_$v.so timeout 300 python scripts/tune.py --only bwd --flags 0,0x20000,0x40000 --dtypes f32 --regimes R
done
