# three alternations of (default build, variant build) on c2, 30 steps; plus MMR_NO_PDL=1 once
tag=$1; var=$2
V=$PWD/mmr_semantic-segmentation_v1_b200/$var
line() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); r=d['roofline']; print(sys.argv[1], round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), round(d['e2e_uint8_frames']['value'],1), round(r['frac'],4), round(r['target_set']['frac'],4), d['clocks']['sm_mhz'])" $1; }
for rep in 1 2 3; do
  python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/${tag}_new$rep.json 2> gpurun_out/${tag}_err.txt; line gpurun_out/${tag}_new$rep.json
  MMR_LIB=$V python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/${tag}_old$rep.json 2>> gpurun_out/${tag}_err.txt; line gpurun_out/${tag}_old$rep.json
done
MMR_NO_PDL=1 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/${tag}_nopdl.json 2>> gpurun_out/${tag}_err.txt; line gpurun_out/${tag}_nopdl.json
