set -x
python -m pytest tests/test_pointwise_head_gpu.py tests/test_resnet_unet_gpu.py -x -q 2>&1 | tail -5
MMR_NO_ARENA_REUSE=1 python -m pytest tests/test_parity_gpu.py -x -q -k "resnet or ResNet or c3 or config3" 2>&1 | tail -5
python scripts/step_breakdown.py 0 c3 2>&1 | head -24
