# last build of the round (r02zh + 16-byte weight packing + 128 KB weight-gradient stage ring): GPU tests, bench lines of every
# config, kernel timeline, step breakdown, smoke, reference arm
tag=${1:-r02zm}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | tee gpurun_out/${tag}_pytest.txt
python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; python -c "
import json; d=json.load(open('gpurun_out/${tag}_bench.json')); r=d['roofline']; print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e_uint8_frames']['value'], r['frac'], r['target_set']['frac'], {k:round(v['frac'],3) for k,v in r['target_set_by_pass'].items()}, r['traffic'], d['clocks'], d['cpu_baseline']['value']); print({k:(v.get('value'), v.get('ms_per_step')) for k,v in d.get('other_configs',{}).items()})"
for c in c3 c4 c5; do python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_$c.json 2> gpurun_out/${tag}_bench_$c.err; python -c "
import json; d=json.load(open('gpurun_out/${tag}_bench_$c.json')); print('$c', d['value'], d['ms_per_step'], d.get('e2e',{}).get('value'), d['roofline']['frac'], d['roofline'].get('target_set',{}).get('frac'))"; done
python scripts/step_timeline.py c2 gpurun_out/${tag}_c2_timeline > /dev/null 2>&1; head -4 gpurun_out/${tag}_c2_timeline.txt
python scripts/step_breakdown.py > gpurun_out/${tag}_c2_step_breakdown.txt 2>&1; head -2 gpurun_out/${tag}_c2_step_breakdown.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -c 300
