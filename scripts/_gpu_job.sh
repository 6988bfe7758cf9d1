for o in 0 1 2; do echo "PCT_HYBRID_ONLY=$o"; PCT_HYBRID_ONLY=$o timeout 300 python scripts/tune.py --only bwd --flags 0x80 --dtypes f32 --regimes R; done
