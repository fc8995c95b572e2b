# like ab_env.sh for another bench config.   usage: ab_env_cfg.sh <tag> <config> <reps> <VAR> <value> ...
tag=$1; cfg=$2; reps=$3; var=$4; shift 4
line() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); r=d['roofline']; print(sys.argv[1], round(d['value'],1), round(d['ms_per_step'],3), round(d.get('e2e',{}).get('value',0),1), round(r['frac'],4), d['clocks']['sm_mhz'])" $1; }
for rep in $(seq 1 $reps); do for v in "$@"; do
  if [ "$v" = "-" ]; then unset $var; else export $var=$v; fi
  python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/${tag}_${cfg}_${v}_$rep.json 2> gpurun_out/${tag}_err.txt || tail -3 gpurun_out/${tag}_err.txt; line gpurun_out/${tag}_${cfg}_${v}_$rep.json
done; done
