set -x
export ONLY=x_1_3.conv1
for cfg in "tx=2 tps=1" "tx=2 tps=3" "tx=4 tps=1" "tx=4 tps=3" "tx=1 tps=3" "tx=4 tps=3 halo_stages=2"; do python scripts/bench_halo.py sw $cfg 2>&1 | grep -v TOTAL; done
export ONLY=x_3_3.conv2
for cfg in "tx=2 tps=3" "tx=4 tps=3" "tx=4 tps=1" "tx=1 tps=3"; do python scripts/bench_halo.py sw $cfg 2>&1 | grep -v TOTAL; done
export ONLY=x_0_1.conv1
for cfg in "tx=2 bn=128" "tx=1 bn=128" "tx=4 bn=128" "tx=2 bn=64 tps=3" "tx=4 bn=64 tps=3"; do python scripts/bench_halo.py sw $cfg 2>&1 | grep -v TOTAL; done
export ONLY=x_0_4
for cfg in "tx=4" "tx=2" "tx=1"; do python scripts/bench_halo.py sw $cfg 2>&1 | grep -v TOTAL; done
