for v in "A=1" "MMR_NO_PDL=1" "MMR_NO_GRAPH=1" "MMR_NO_GRAPH=1 MMR_NO_PDL=1"; do
  echo "== $v"
  for i in 1 2; do env $v python bench.py --steps 20 --warmup 3 --no-extra --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"; done
done
