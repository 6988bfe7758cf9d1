for f in 0 0x20; do
echo "== PCT_MSDA_BWD_FLAGS=$f"
PCT_MSDA_BWD_FLAGS=$f timeout 300 python scripts/bench_encoder.py --mode train --batch 16 2>&1 | tail -1 | cut -c1-400
PCT_MSDA_BWD_FLAGS=$f timeout 300 python scripts/bench_encoder.py --mode train --pyramid bbbc --batch 32 2>&1 | tail -1 | cut -c1-400
done
