timeout 900 python bench.py > gpurun_out/bench_r02_f32.json 2> gpurun_out/bench_r02_f32.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02_f32.json'))
print(round(d['value']/1e6,2), round(d['ms_per_step'],4), d['roofline']['frac'], d['forward']['ms'], d['backward']['ms'], d['e2e']['value'], d['cpu_baseline']['value'], d['gpu_launches'])
for k in ('cfg4_bbbc_encoder6_train_b32','cfg2_decoder_infer_b8','cfg5_encoder6_train'):
    v=d['sub_records'][k]; print(k, round(v['ms_per_step'],2), round(v['images_per_s'],1))
PY
timeout 300 python scripts/bench_encoder.py --mode infer --batch 8 2>&1 | tail -1 | cut -c1-300
