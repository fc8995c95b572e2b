export STATS=1
for L in x_3_3.conv2 x_1_3.conv1 layer1.conv x_2_3.conv1; do
for K in fprop dgrad; do
  echo -n "auto    "; ONLY=$L KINDS=$K python scripts/bench_halo.py d0 2>&1 | grep -E "^$L" | cut -c1-100
  echo -n "direct  "; ONLY=$L KINDS=$K python scripts/bench_halo.py d1 direct=1 2>&1 | grep -E "^$L" | cut -c1-100
done; done
