import os, sys, torch
sys.path.insert(0, os.getcwd())
from tests.helpers import install_engine_masks, model_pair, rel, synthetic_batch
from oracle.losses import mixed_loss
from mmrseg_b200.losses import DiceCrossEntropyLoss
ref, net = model_pair(10, "resnet34")
n, hw = int(sys.argv[1]), int(sys.argv[2])
x, y = synthetic_batch(n, 10, hw, hw)
ref.train(); net.train()
got = net(x.cuda())
loss = DiceCrossEntropyLoss(0.5)(got, y.cuda()); loss.backward(); torch.cuda.synchronize()
eng = list(net._engines.values())[0]
install_engine_masks(ref, eng)
want = ref(x); loss_ref = mixed_loss(want, y, 0.5); loss_ref.backward()
print("logits rel", rel(got.detach().cpu(), want.detach()), "loss", loss.item(), loss_ref.item())
rp = dict(ref.named_parameters())
rows = []
for name, p in net.named_parameters():
    g, r = p.grad.cpu(), rp[name].grad
    cos = torch.nn.functional.cosine_similarity(g.flatten(), r.flatten(), dim=0).item()
    rows.append((rel(g, r), cos, name))
rows.sort(reverse=True)
for r in rows[:4]: print("%.4f %.4f %s" % r)
print("halo units:", sum(1 for u in eng.units if u.get("halo")), "of", len(eng.units))
