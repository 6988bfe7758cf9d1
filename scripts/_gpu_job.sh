timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 > gpurun_out/bench_r02_f32_n8.json 2> gpurun_out/bench_r02_f32_n8.err
tail -2 gpurun_out/bench_r02_f32_n8.err
python -c "
import json
d=json.load(open('gpurun_out/bench_r02_f32_n8.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value']); print(d['sub_records']['cfg5_encoder6_train'])"
