for v in on off; do
  if [ $v = off ]; then export MMR_NO_PDL=1; else unset MMR_NO_PDL; fi
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('PDL $v', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])"
done
