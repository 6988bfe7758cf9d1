timeout 600 ncu --set full --clock-control none -k regex:"msda_fwd" -c 4 -o gpurun_out/r02_fwd_staged python scripts/tune.py --only fwd --flags 0,0x4 --dtypes f32 --regimes R --iters 1 --warm 1 > gpurun_out/r02_fwd_staged_ncu.log 2>&1
( timeout 300 python scripts/tune.py --only fwd --flags 0,0x4 --dtypes f32 --regimes R,U
  timeout 300 python scripts/tune.py --only fwd --flags 0,0x4 --dtypes f32 --regimes R --pyramid BBBC_520x696 --batch 32
  timeout 300 python scripts/tune.py --only fwd --flags 0,0x4 --dtypes f32 --regimes R --batch 4 ) > gpurun_out/r02_fwd_staged_tiles.log 2>&1
