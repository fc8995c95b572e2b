# N=8 attribution of the data-parallel overhead (run with gpurun --gpus 8)
run() { echo "== $*"; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 --no-extra --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"; }
run A=1
run MMR_DDP_NO_COMM=1
run NCCL_ALGO=NVLS
run NCCL_ALGO=Ring
nvidia-smi --query-gpu=index,clocks.sm,power.draw,temperature.gpu --format=csv
