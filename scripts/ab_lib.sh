# A/B of two builds of the library on ONE box: alternating bench runs (c2 by default), then the GPU tests on the default build.
# usage: ab_lib.sh <tag> <variant .so name> [config ...]
tag=$1; var=$2; shift 2; cfgs=${@:-c2}
V=$PWD/mmr_semantic-segmentation_v1_b200/$var
line() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); r=d['roofline']; print(sys.argv[1], d['value'], d['ms_per_step'], d.get('e2e',{}).get('value'), r['frac'], r.get('target_set',{}).get('frac'), d['clocks']['sm_mhz'])" $1; }
for c in $cfgs; do for rep in 1 2; do
  python bench.py --config $c --steps 20 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/${tag}_${c}_new$rep.json 2> gpurun_out/${tag}_err.txt; line gpurun_out/${tag}_${c}_new$rep.json
  MMR_LIB=$V python bench.py --config $c --steps 20 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/${tag}_${c}_old$rep.json 2>> gpurun_out/${tag}_err.txt; line gpurun_out/${tag}_${c}_old$rep.json
done; done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
