# several environment settings on ONE box, alternating: each argument is "name:VAR=value[,VAR=value]" ("name:" = defaults)
tag=$1; reps=$2; shift 2
line() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); r=d['roofline']; print(sys.argv[1], round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), round(r['frac'],4), round(r['target_set']['frac'],4), d['clocks']['sm_mhz'])" $1; }
for rep in $(seq 1 $reps); do for spec in "$@"; do
  name=${spec%%:*}; envs=${spec#*:}
  ( IFS=,; for kv in $envs; do [ -n "$kv" ] && export "$kv"; done
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/${tag}_${name}_$rep.json 2> gpurun_out/${tag}_err.txt || tail -3 gpurun_out/${tag}_err.txt )
  line gpurun_out/${tag}_${name}_$rep.json
done; done
