set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 > gpurun_out/r02u_bench.json 2> gpurun_out/r02u_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r02u_bench.json')); r=d['roofline']; print(d['value'], d['ms_per_step'], d['e2e']['value'], r['frac'], r['target_set']['frac'], {k:round(v['frac'],3) for k,v in r['target_set_by_pass'].items()}, d['clocks'])"
for c in c3 c4 c5; do python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02u_bench_$c.json 2> gpurun_out/r02u_bench_$c.err; python -c "
import json; d=json.load(open('gpurun_out/r02u_bench_$c.json')); print('$c', d['value'], d['ms_per_step'], d.get('e2e',{}).get('value'), d['roofline']['frac'])"; done
timeout 200 python scripts/layer_table.py r02u > gpurun_out/r02u_layers.log 2>&1; head -4 gpurun_out/r02u_layers.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -c 400
