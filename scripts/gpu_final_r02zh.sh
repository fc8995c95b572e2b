# last build of the round (tensor-map prefetch, median-of-seven roofline sampling): GPU tests, bench lines of every config, smoke
tag=${1:-r02zh}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | tee gpurun_out/${tag}_pytest.txt
python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; python -c "
import json; d=json.load(open('gpurun_out/${tag}_bench.json')); r=d['roofline']; print(d['value'], d['ms_per_step'], d['e2e']['value'], r['frac'], r['target_set']['frac'], {k:round(v['frac'],3) for k,v in r['target_set_by_pass'].items()}, r['traffic'], d['clocks']); print({k:(v.get('value'), v.get('ms_per_step')) for k,v in d.get('other_configs',{}).items()})"
for c in c3 c4 c5; do python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_$c.json 2> gpurun_out/${tag}_bench_$c.err; python -c "
import json; d=json.load(open('gpurun_out/${tag}_bench_$c.json')); print('$c', d['value'], d['ms_per_step'], d.get('e2e',{}).get('value'), d['roofline']['frac'], d['roofline'].get('target_set',{}).get('frac'))"; done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
