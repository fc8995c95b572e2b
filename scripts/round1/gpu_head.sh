export KINDS=fprop ONLY=head10
for D in 0 1 4 8 12 5 13; do echo -n "dbg=$D "; MMR_HALO_DBG=$D python scripts/bench_halo.py head 2>&1 | grep -E "^head" | cut -c1-110; done
