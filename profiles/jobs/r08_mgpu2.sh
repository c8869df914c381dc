python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_check.py 2>&1 | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r08_bench2.json 2> gpurun_out/r08_bench2.err
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/r08_bench2.json') if l.startswith('{')][-1])
print('N=2 headline', round(d['ms_per_step'],3), round(d['value']), round(d['e2e']['value']), d.get('gradient_exchange'))
for e in d.get('extra_configs', []): print(e.get('name'), round(e.get('ms_per_step',0),3), round(e.get('value',0)))
P
