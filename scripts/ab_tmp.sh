set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 > gpurun_out/r02o_bench.json 2> gpurun_out/r02o_bench.err; tail -c 600 gpurun_out/r02o_bench.json
for c in c3 c4 c5; do python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02o_bench_$c.json 2> gpurun_out/r02o_bench_$c.err; python -c "
import json; d=json.load(open('gpurun_out/r02o_bench_$c.json')); print('$c', d['value'], d['ms_per_step'], d.get('e2e',{}).get('value'), d['roofline']['frac'])"; done
