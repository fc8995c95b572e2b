# the build the round ends with (r02zm + logits returned as a view for one-head models): GPU tests, the default bench line
# (c2 + short c3 / c4 / c5 lines), smoke
tag=${1:-r02zp}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | tee gpurun_out/${tag}_pytest.txt
python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; python -c "
import json; d=json.load(open('gpurun_out/${tag}_bench.json')); r=d['roofline']; print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e_uint8_frames']['value'], r['frac'], r['target_set']['frac'], {k:round(v['frac'],3) for k,v in r['target_set_by_pass'].items()}, d['clocks'], d['cpu_baseline']['value']); print({k:(v.get('value'), v.get('ms_per_step')) for k,v in d.get('other_configs',{}).items()})"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
