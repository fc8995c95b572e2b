export STATS=1 KINDS=fprop
for L in x_1_3.conv1 x_3_3.conv2; do
for F in "rph=1 tx=4" "rph=1 tx=2" "rph=2 tx=1" "rph=2 tx=2" "rph=2 tx=2 acc_bufs=1"; do
for D in 0 3 12; do
  echo -n "$F dbg=$D  "; MMR_HALO_DBG=$D ONLY=$L python scripts/bench_halo.py sw $F 2>&1 | grep -E "^$L" | cut -c1-110
done; done; done
