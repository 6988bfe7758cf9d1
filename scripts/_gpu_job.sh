for b in 2 4 8 16; do timeout 300 python scripts/tune.py --only bwd --flags 0x20,0x40,0x32000040,0x48000040,0x64000040 --dtypes f32 --regimes R --batch $b; done
