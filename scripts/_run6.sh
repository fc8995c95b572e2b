python -m pytest tests/test_loss_metric_gpu.py -x -q 2>&1 | tail -2
python - <<'P'
import torch, time
from mmrseg_b200.losses import DiceCrossEntropyLoss
crit = DiceCrossEntropyLoss(0.5)
for (n,c) in ((32,10),(16,2)):
    x = torch.randn(n,c,512,512,device="cuda",requires_grad=True); y = torch.randint(0,c,(n,512,512),device="cuda")
    for it in range(3):
        l = crit(x,y); l.backward()
    torch.cuda.synchronize()
    e0,e1,e2 = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf=tb=0
    for it in range(5):
        x.grad=None
        e0.record(); l = crit(x,y); e1.record(); l.backward(); e2.record(); torch.cuda.synchronize()
        tf+=e0.elapsed_time(e1)/5; tb+=e1.elapsed_time(e2)/5
    print("loss N=%d C=%d: fwd %.3f ms bwd %.3f ms (through autograd)"%(n,c,tf,tb))
P
for c in c3 c2; do python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$c', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])"; done
