timeout 900 python -m pytest tests/test_conv_halo_gpu.py tests/test_conv_bwd_gpu.py tests/test_blocks_gpu.py -x -q 2>&1 | tail -2
export ONLY=x_0_ KINDS=fprop,dgrad
python scripts/bench_halo.py t 2>&1 | grep "x_0_"
