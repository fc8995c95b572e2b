# 8 GPUs of one box, final build: the bench under torchrun (c2 weak scaling + the c4 line: global batch 256, deep supervision)
tag=${1:-r02y}
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/${tag}_bench_8gpu.json 2> gpurun_out/${tag}_bench_8gpu.err
python -c "
import json
d=json.load(open('gpurun_out/${tag}_bench_8gpu.json'))
print('8 gpu', d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])
print({k:(v.get('value'), v.get('ms_per_step'), v.get('error')) for k,v in d.get('other_configs',{}).items()})"
tail -3 gpurun_out/${tag}_bench_8gpu.err
