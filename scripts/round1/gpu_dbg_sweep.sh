# which stage bounds a layer: time it with parts of the kernel disabled (MMR_HALO_DBG bits: 1 no MMA, 2 no epilogue, 4 no halo TMA, 8 no weight TMA)
export STATS=1
for L in x_0_4.conv2 x_0_4.conv1 x_3_3.conv2 x_1_3.conv1 layer1.conv; do
for K in fprop dgrad; do
for D in 0 1 2 3 4 8 12 6 7 15; do
  echo -n "dbg=$D  "; MMR_HALO_DBG=$D ONLY=$L KINDS=$K python scripts/bench_halo.py dbg 2>&1 | grep -E "^$L" | cut -c1-60
done; done; done
