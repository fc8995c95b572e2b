for k in 4 16 64; do echo "== MMR_BN_ITERS_PER_CTA=$k"; MMR_BN_ITERS_PER_CTA=$k python scripts/reduce_table.py 2>&1 | grep -v "^\[" | awk '{print}' | tail -34 | awk '/16x16|32x32|64x64|total/'; done
