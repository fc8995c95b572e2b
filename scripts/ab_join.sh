# weight-gradient join depth (MMR_WGRAD_JOIN) sweep on one box
tag=$1; shift
line() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); r=d['roofline']; print(sys.argv[1], round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), round(r['frac'],4), d['clocks']['sm_mhz'], d.get('mem_gb'))" $1; }
for c in c2 c3; do for k in 2 3 4 6 10 100 2 6; do
  MMR_WGRAD_JOIN=$k python bench.py --config $c --steps 20 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/${tag}_${c}_k$k.json 2> gpurun_out/${tag}_err.txt || tail -3 gpurun_out/${tag}_err.txt; line gpurun_out/${tag}_${c}_k$k.json
done; done
MMR_WGRAD_JOIN=6 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
