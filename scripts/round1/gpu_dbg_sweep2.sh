export STATS=1
for L in x_0_4.conv2 x_1_3.conv1 x_3_3.conv2; do
for K in fprop dgrad; do
for D in 0 1 3 12 15; do
  echo -n "dbg=$D  "; MMR_HALO_DBG=$D ONLY=$L KINDS=$K python scripts/bench_halo.py dbg 2>&1 | grep -E "^$L" | cut -c1-60
done; done; done
python scripts/bench_halo.py spin 2>&1 | tail -4
python scripts/step_breakdown.py 2>&1 | grep -E "step |graph|eager"
