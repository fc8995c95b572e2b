# A/B of an engine environment switch on one box: bash scripts/round1/gpu_ab_env.sh MMR_NO_FUSED_BWD_REDUCE
for rep in 1 2; do for v in on off; do
  if [ $v = off ]; then export $1=1; else unset $1; fi
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 feature $v', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])"
done; done
