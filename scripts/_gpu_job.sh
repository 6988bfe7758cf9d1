timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python scripts/tune.py --flags 0 --dtypes f32,bf16 --regimes R
timeout 300 python scripts/tune.py --flags 0 --dtypes f32 --regimes R --d-model 128
