export KINDS=dgrad
for L in x_2_3.conv1 x_0_2.conv1; do
export ONLY=$L
for cfg in "bn=64 tx=4" "bn=64 tx=2" "bn=192 tx=2" "bn=192 tx=1" "bn=160 tx=1" "bn=128 tx=2" "bn=96 tx=2" "bn=96 tx=4"; do python scripts/bench_halo.py sw $cfg 2>&1 | grep -v TOTAL; done
done
export KINDS=fprop ONLY=x_0_3.conv1
for cfg in "tx=4 tps=1" "tx=4 tps=3" "tx=2 tps=3"; do python scripts/bench_halo.py sw $cfg 2>&1 | grep -v TOTAL; done
