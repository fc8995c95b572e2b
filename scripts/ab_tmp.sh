timeout 600 python -m pytest tests/test_conv_halo_gpu.py -x -q 2>&1 | tail -2
export ONLY=x_0_4 
for cfg in "loader=0" "loader=1" "loader=1 tx=4 rph=2" "loader=1 tx=4 rph=1" "loader=1 tx=2 rph=4"; do echo "$cfg:"; KINDS=fprop,dgrad python scripts/bench_halo.py t $cfg 2>&1 | grep x_0_4; done
echo "no-epilogue / no-MMA views of loader=1 tx=4 rph=2 on conv2 fprop:"
for dbg in 2 1 5; do echo "dbg=$dbg $(ONLY=x_0_4.conv2 KINDS=fprop MMR_HALO_DBG=$dbg python scripts/bench_halo.py t loader=1 tx=4 rph=2 2>&1 | grep x_0_4)"; done
