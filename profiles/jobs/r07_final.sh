timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py > gpurun_out/r07_bench1.json 2> gpurun_out/r07_bench1.err
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/r07_bench1.json') if l.startswith('{')][-1])
print('headline', round(d['ms_per_step'],3), round(d['value']), round(d['e2e']['value']), round(d['roofline']['frac'],3), d['gpu_launches'], d['clocks'])
for e in d.get('extra_configs', []): print(e.get('name'), round(e.get('ms_per_step',0),3), round(e.get('value',0)))
print(d.get('hbm_stages'))
P
