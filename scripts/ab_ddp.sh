# N=2 A/B of the DDP knobs (run with gpurun --gpus 2)
run() { echo "== $*"; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-extra --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])"; }
python bench.py --steps 20 --warmup 3 --no-extra --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N=1', d['value'], d['ms_per_step'])"
run A=1
run MMR_DDP_BUCKET_MB=16
run MMR_DDP_BUCKET_MB=32
run MMR_DDP_BUCKET_MB=64
run MMR_DDP_BUCKET_MB=16 MMR_DDP_TAIL_MB=4
run NCCL_MAX_NCHANNELS=4
run NCCL_MAX_NCHANNELS=2 MMR_DDP_BUCKET_MB=16
run A=1
