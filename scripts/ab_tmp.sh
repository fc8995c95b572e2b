timeout 900 python -m pytest tests/test_model_gpu.py tests/test_parity_gpu.py -x -q 2>&1 | tail -2
for v in "MMR_WGRAD_LATE=0 MMR_WGRAD_SMS=148" "MMR_WGRAD_LATE=1 MMR_WGRAD_SMS=148" "MMR_WGRAD_LATE=0 MMR_WGRAD_SMS=111" "MMR_WGRAD_LATE=1 MMR_WGRAD_SMS=111" "MMR_WGRAD_LATE=1 MMR_WGRAD_SMS=111 MMR_FUSED_BWD_CH=" "MMR_WGRAD_LATE=1 MMR_WGRAD_SMS=96"; do
  echo "== $v"
  env $v python bench.py --steps 20 --warmup 3 --no-extra --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'])"
done
