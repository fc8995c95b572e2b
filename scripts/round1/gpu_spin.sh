export STATS=1
for L in layer1.conv layer3.conv x_1_3.conv1; do for K in fprop dgrad; do for D in 0 15; do
  echo -n "dbg=$D  "; MMR_HALO_DBG=$D ONLY=$L KINDS=$K python scripts/bench_halo.py spin 2>&1 | grep -E "^$L" | cut -c1-60
done; done; done
python scripts/bench_halo.py spin 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'])"
