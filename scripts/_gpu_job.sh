set -x
O=gpurun_out
timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/r02f_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r02f_smoke.log 2>&1
timeout 900 python bench.py > $O/bench_r02_f32.json 2> $O/bench_r02_f32.err
timeout 600 python bench.py --dtype bf16 --no-sub-records > $O/bench_r02_bf16.json 2> $O/bench_r02_bf16.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_r02_reference.json 2> $O/bench_r02_reference.err
( timeout 300 python scripts/tune.py --flags 0x20,0 --dtypes f32,bf16 --regimes R,U
  timeout 300 python scripts/tune.py --flags 0x20,0 --dtypes f32,bf16 --regimes R --d-model 128
  timeout 300 python scripts/tune.py --flags 0x20,0 --dtypes f32 --regimes R --d-model 512
  timeout 300 python scripts/tune.py --flags 0x20,0 --dtypes f32,bf16 --regimes R --pyramid BBBC_520x696 --batch 32
  for b in 1 2 4 8; do timeout 300 python scripts/tune.py --only bwd --flags 0x20,0 --dtypes f32 --regimes R --batch $b; done ) > $O/r02_tune_final.log 2>&1
timeout 600 python scripts/bench_legacy.py > $O/r02_legacy_vs_new_f32_d32.json 2>$O/legacy.err
timeout 600 python scripts/bench_legacy.py --d-model 128 > $O/r02_legacy_vs_new_f32_d16.json 2>>$O/legacy.err
timeout 600 python scripts/bench_legacy.py --regime U > $O/r02_legacy_vs_new_f32_d32_uniform.json 2>>$O/legacy.err
timeout 300 python scripts/bench_fused.py 16 > $O/r02_fused_vs_unfused_op.json 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"msda_(fwd|bwd)" -c 4 -o $O/r02_final python scripts/tune.py --flags 0 --dtypes f32 --regimes R --iters 1 --warm 1 > $O/r02_final_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_ncu_launch_list.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-sub-records > $O/r02_launch_list_run.log 2>&1
