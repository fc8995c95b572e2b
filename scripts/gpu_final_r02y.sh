# final build of round 2: GPU tests, bench lines of every config, kernel timelines (CUPTI), launch list + per-kernel
# DRAM traffic of a c2 step (ncu, after the same command exited 0 without ncu), one full ncu capture, smoke, reference arm
tag=${1:-r02y}
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/${tag}_pytest.txt
python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; python -c "
import json; d=json.load(open('gpurun_out/${tag}_bench.json')); r=d['roofline']; print(d['value'], d['ms_per_step'], d['e2e']['value'], r['frac'], r['target_set']['frac'], {k:round(v['frac'],3) for k,v in r['target_set_by_pass'].items()}, d['clocks']); print({k:(v.get('value'), v.get('ms_per_step')) for k,v in d.get('other_configs',{}).items()})"
for c in c3 c4 c5; do python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_$c.json 2> gpurun_out/${tag}_bench_$c.err; python -c "
import json; d=json.load(open('gpurun_out/${tag}_bench_$c.json')); print('$c', d['value'], d['ms_per_step'], d.get('e2e',{}).get('value'), d['roofline']['frac'])"; done
python scripts/step_timeline.py c2 gpurun_out/${tag}_c2_timeline > /dev/null 2> gpurun_out/${tag}_timeline.err; head -4 gpurun_out/${tag}_c2_timeline.txt
python scripts/step_timeline.py c3 gpurun_out/${tag}_c3_timeline > /dev/null 2>> gpurun_out/${tag}_timeline.err; head -4 gpurun_out/${tag}_c3_timeline.txt
python scripts/step_breakdown.py > gpurun_out/${tag}_c2_step_breakdown.txt 2>&1; head -3 gpurun_out/${tag}_c2_step_breakdown.txt
MMR_BENCH_LAYERS=gpurun_out/${tag}_layers_instep.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra > /dev/null 2>&1
MMR_NO_GRAPH=1 python scripts/profile_step.py 16 4 > gpurun_out/${tag}_c2_plain.log 2>&1 && \
MMR_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 640 -c 760 --csv --log-file gpurun_out/${tag}_c2_launches.csv python scripts/profile_step.py 16 4 > gpurun_out/${tag}_c2_ncu.log 2>&1
grep 'gpu__time_duration\|^"ID"' gpurun_out/${tag}_c2_launches.csv > gpurun_out/${tag}_c2_launch_times.csv
python scripts/launch_summary.py gpurun_out/${tag}_c2_launch_times.csv > gpurun_out/${tag}_c2_launch_summary.txt 2>&1; head -8 gpurun_out/${tag}_c2_launch_summary.txt
python scripts/conv_traffic.py gpurun_out/${tag}_c2_launches.csv gpurun_out/${tag}_c2_step_traffic.txt | tail -3
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad_kx_kernel -s 12 -c 1 -o gpurun_out/${tag}_wgrad_kx python scripts/profile_step.py 16 2 > gpurun_out/${tag}_ncu_full.log 2>&1; tail -2 gpurun_out/${tag}_ncu_full.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -c 600
