# A/B of an environment knob on ONE box: alternating c2 bench runs.   usage: ab_env.sh <tag> <reps> <VAR> <value> ...   ("-" = unset)
tag=$1; reps=$2; var=$3; shift 3
line() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); r=d['roofline']; print(sys.argv[1], round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), round(r['frac'],4), round(r['target_set']['frac'],4), round(r['by_pass']['wgrad']['frac'],3), d['clocks']['sm_mhz'])" $1; }
for rep in $(seq 1 $reps); do for v in "$@"; do
  if [ "$v" = "-" ]; then unset $var; else export $var=$v; fi
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/${tag}_${var}_${v}_$rep.json 2> gpurun_out/${tag}_err.txt || tail -3 gpurun_out/${tag}_err.txt; line gpurun_out/${tag}_${var}_${v}_$rep.json
done; done
