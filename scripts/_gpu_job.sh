for v in noopen nosweep; do
  echo "== variant $v"
  PCTRANS_MSDA_LIB=pctrans_b200/_build/variants/libpctrans_msda_$v.so timeout 300 python scripts/tune.py --only bwd --flags 0x40 --dtypes f32 --regimes R
done
