# 2 GPUs of one box, final build: torchrun DDP tests, then the bench under torchrun (c2 + the c4 line)
tag=${1:-r02y}
set -x
timeout 600 python -m pytest tests/test_ddp_gpu.py -x -q 2>&1 | tail -3 | tee gpurun_out/${tag}_pytest_ddp.txt
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/${tag}_bench_1of2.json 2> /dev/null
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/${tag}_bench_2gpu.json 2> gpurun_out/${tag}_bench_2gpu.err
python -c "
import json
a=json.load(open('gpurun_out/${tag}_bench_1of2.json')); d=json.load(open('gpurun_out/${tag}_bench_2gpu.json'))
print('1 gpu', a['value'], a['ms_per_step']); print('2 gpu', d['value'], d['ms_per_step'], d['e2e']['value'], 'eff', d['value']/(2*a['value']))
print({k:(v.get('value'), v.get('ms_per_step'), v.get('error')) for k,v in d.get('other_configs',{}).items()})"
