python -m pytest tests/test_conv_halo_gpu.py tests/test_resnet_unet_gpu.py -x -q 2>&1 | tail -2
echo "== c3 default (8 affine epilogue warps)"; python scripts/step_breakdown.py 0 c3 2>&1 | grep -E "^step|halo_conv_plan_run"
echo "== c3 MMR_AFFINE_EPI_WARPS=4"; MMR_AFFINE_EPI_WARPS=4 python scripts/step_breakdown.py 0 c3 2>&1 | grep -E "^step|halo_conv_plan_run"
echo "== c5 default"; python bench.py --config c5 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'])"
echo "== c5 4 warps"; MMR_AFFINE_EPI_WARPS=4 python bench.py --config c5 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'])"
