mkdir -p gpurun_out
timeout 300 python scripts/bench_legacy.py > gpurun_out/r02_legacy_vs_new_f32_d32.json 2>gpurun_out/legacy.err; tail -1 gpurun_out/r02_legacy_vs_new_f32_d32.json | cut -c1-600
timeout 300 python scripts/bench_legacy.py --d-model 128 > gpurun_out/r02_legacy_vs_new_f32_d16.json 2>>gpurun_out/legacy.err; tail -1 gpurun_out/r02_legacy_vs_new_f32_d16.json | cut -c1-600
timeout 300 python scripts/bench_legacy.py --regime U > gpurun_out/r02_legacy_vs_new_f32_d32_uniform.json 2>>gpurun_out/legacy.err; tail -1 gpurun_out/r02_legacy_vs_new_f32_d32_uniform.json | cut -c1-600
tail -3 gpurun_out/legacy.err
timeout 600 python bench.py > gpurun_out/bench_r02_f32.json 2> gpurun_out/bench_r02_f32.err; echo bench rc=$?
timeout 300 python bench.py --dtype bf16 --no-sub-records > gpurun_out/bench_r02_bf16.json 2>> gpurun_out/bench_r02_f32.err; echo bench bf16 rc=$?
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_r02_reference.json 2>> gpurun_out/bench_r02_f32.err; echo ref rc=$?
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-sub-records > gpurun_out/plain2.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launch_list.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-sub-records > gpurun_out/ncu2.log 2>&1; echo ncu rc=$?
