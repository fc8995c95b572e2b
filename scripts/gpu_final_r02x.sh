# last build of round 2: GPU tests, bench lines of all configs, c2 / c3 launch lists and step breakdowns, per-layer tables
tag=${1:-r02x}
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; python -c "
import json; d=json.load(open('gpurun_out/${tag}_bench.json')); r=d['roofline']; print(d['value'], d['ms_per_step'], d['e2e']['value'], r['frac'], r['target_set']['frac'], {k:round(v['frac'],3) for k,v in r['target_set_by_pass'].items()}, d['clocks']); print({k:(v.get('value'), v.get('ms_per_step')) for k,v in d.get('other_configs',{}).items()})"
for c in c3 c4 c5; do python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_$c.json 2> gpurun_out/${tag}_bench_$c.err; python -c "
import json; d=json.load(open('gpurun_out/${tag}_bench_$c.json')); print('$c', d['value'], d['ms_per_step'], d.get('e2e',{}).get('value'), d['roofline']['frac'])"; done
python scripts/step_breakdown.py > gpurun_out/${tag}_c2_step_breakdown.txt 2>&1; head -3 gpurun_out/${tag}_c2_step_breakdown.txt
python scripts/step_breakdown.py 0 c3 > gpurun_out/${tag}_c3_step_breakdown.txt 2>&1; head -3 gpurun_out/${tag}_c3_step_breakdown.txt
timeout 200 python scripts/layer_table.py ${tag} > gpurun_out/${tag}_layers.log 2>&1; head -4 gpurun_out/${tag}_layers.log
timeout 200 python scripts/layer_table.py ${tag}_c3 0 c3 > gpurun_out/${tag}_c3_layers.log 2>&1; head -4 gpurun_out/${tag}_c3_layers.log
MMR_NO_GRAPH=1 python scripts/profile_step.py 0 3 c3 > gpurun_out/${tag}_c3_plain.log 2>&1 && \
MMR_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${tag}_c3_launches.csv python scripts/profile_step.py 0 3 c3 > gpurun_out/${tag}_c3_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/${tag}_c3_launches.csv 1 > gpurun_out/${tag}_c3_launch_summary.txt 2>&1; head -12 gpurun_out/${tag}_c3_launch_summary.txt
MMR_NO_GRAPH=1 python scripts/profile_step.py 16 4 > gpurun_out/${tag}_c2_plain.log 2>&1 && \
MMR_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 640 -c 760 --csv --log-file gpurun_out/${tag}_c2_launches.csv python scripts/profile_step.py 16 4 > gpurun_out/${tag}_c2_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/${tag}_c2_launches.csv > gpurun_out/${tag}_c2_launch_summary.txt 2>&1; head -8 gpurun_out/${tag}_c2_launch_summary.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pointwise_head_bwd_kernel -c 1 -o gpurun_out/${tag}_pointwise_bwd python scripts/profile_step.py 0 2 c3 > gpurun_out/${tag}_ncu_pw.log 2>&1; tail -2 gpurun_out/${tag}_ncu_pw.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -c 400
