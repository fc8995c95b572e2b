timeout 900 python -m pytest tests/test_conv_halo_gpu.py tests/test_conv_bwd_gpu.py tests/test_blocks_gpu.py tests/test_model_gpu.py tests/test_parity_gpu.py tests/test_data_gpu.py -x -q 2>&1 | tail -4
MMR_HALO_DBG=16 python scripts/halo_trace.py layer1 x_3_3.conv2 2>&1 | head -40
python bench.py --steps 20 --warmup 3 --no-extra --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['target_set']['frac'])"
