python -m pytest tests/test_resnet_unet_gpu.py tests/test_elementwise_gpu.py -x -q 2>&1 | tail -2
MMR_NO_ARENA_REUSE=1 python -m pytest tests/test_parity_gpu.py -x -q 2>&1 | tail -2
python scripts/bench_bilinear.py 2>&1 | tail -7
python scripts/step_breakdown.py 0 c3 2>&1 | grep -E "^step|bilinear|pointwise"
python scripts/step_breakdown.py 32 c4 2>&1 | grep -E "^step|nearest|sumpool"
