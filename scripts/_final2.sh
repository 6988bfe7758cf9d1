set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
( timeout 200 python scripts/tune.py --dtypes f32,bf16 --regimes R,U --iters 40
  timeout 200 python scripts/tune.py --dtypes f32,bf16 --regimes R --iters 40 --d-model 128
  timeout 200 python scripts/tune.py --dtypes f32,bf16 --regimes R --iters 30 --batch 32 --pyramid BBBC_520x696 ) > gpurun_out/tune_final.log 2>&1
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/plain_short.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:msda -c 12 -f -o gpurun_out/final_r01 python scripts/tune.py --iters 1 --warm 1 --dtypes f32,bf16 > gpurun_out/ncu_full.log 2>&1
timeout 600 python bench.py > gpurun_out/bench_r01_f32.json 2> gpurun_out/bench_err.log
timeout 600 python bench.py --dtype bf16 > gpurun_out/bench_r01_bf16.json 2>> gpurun_out/bench_err.log
timeout 600 python bench.py --impl reference > gpurun_out/bench_r01_reference.json 2>> gpurun_out/bench_err.log
timeout 300 python scripts/bench_legacy.py > gpurun_out/legacy_f32_d32.json 2> gpurun_out/legacy_err.log
cat gpurun_out/bench_r01_f32.json
