"""The reference's literal call sites run against the drop-in (VERDICT round 1 "weak #3",
ADVICE round 1): stock constructor arguments, `.half()` + channels_last inference, `clip_grad_norm_(parameters,
12)`, fresh output tensors, eval after a weight update, optimiser param-group subsets and state_dict resume."""
import warnings

import pytest
import torch

from tests.helpers import model_pair, rel, synthetic_batch

pytestmark = pytest.mark.gpu


def test_stock_constructor_lines_do_not_raise():
    """SU/ModelTraining.py:248-253 and ED/Main_MMR_SegModel.py:589 verbatim (smp = mmrseg_b200.models)."""
    import mmrseg_b200.models as smp
    num_classes = 10
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")          # no torch-hub cache on the box: random init + warning
        model = smp.UnetPlusPlus(
            encoder_name="resnet18",
            encoder_weights="imagenet",
            in_channels=3,
            classes=num_classes,
        )
        cfg = {"arch": "UnetPlusPlus", "encoder_name": "resnet34", "encoder_weights": "imagenet", "in_channels": 3,
               "classes": num_classes}
        model2 = smp.create_model(**cfg)
        from mmrseg_b200.models import ResNetUNet
        model3 = ResNetUNet(n_class=num_classes, resnet_model=18)
    assert sum(p.numel() for p in model.parameters()) == 15_971_754
    assert model2.encoder_name == "resnet34" and model3.n_class == 10
    with pytest.raises(KeyError):
        smp.UnetPlusPlus(encoder_name="resnet18", encoder_weights="ssl")


def test_imagenet_weights_come_from_the_local_cache(tmp_path, monkeypatch):
    import torchvision
    import mmrseg_b200.models as smp
    torch.manual_seed(1)
    tv = torchvision.models.resnet18(weights=None)
    torch.save(tv.state_dict(), tmp_path / "resnet18-f37072fd.pth")
    monkeypatch.setenv("MMRSEG_PRETRAINED_DIR", str(tmp_path))
    with warnings.catch_warnings():
        warnings.simplefilter("error")           # found: no warning
        model = smp.UnetPlusPlus(encoder_name="resnet18", encoder_weights="imagenet", in_channels=3, classes=2)
    sd = tv.state_dict()
    for k, v in model.encoder.state_dict().items():
        assert torch.equal(v, sd[k]), k


def test_half_channels_last_inference_lines():
    """ED/Main_MMR_SegModel.py:1240-1267: model.eval(); .to(memory_format=channels_last); .half(); frames in
    fp16 channels_last under inference_mode."""
    ref, net = model_pair(10)
    x, _ = synthetic_batch(2, 10, 64, 96)
    net.eval()
    with torch.no_grad():
        want32 = net(x.cuda())
    net = net.to(memory_format=torch.channels_last)
    net.half()
    assert all(p.dtype == torch.float32 for p in net.parameters())     # masters stay fp32
    with torch.inference_mode():
        frames = x.cuda().to(dtype=torch.float16, memory_format=torch.channels_last)
        preds = net(frames)
    assert preds.dtype == torch.float16 and preds.shape == (2, 10, 64, 96)
    assert preds.is_contiguous(memory_format=torch.channels_last)
    # the differences: fp16 rounding of the frames and of the returned logits.  The first is tiny (5e-4) but
    # moves bf16 roundings inside the network, which decorrelates them layer by layer (tests/teacher.py): the
    # bound is the bf16 noise level of an eval forward, and the masks agree wherever the margin is clear
    assert rel(preds.float(), want32) <= 3e-2
    top2 = want32.topk(2, dim=1).values
    safe = (top2[:, 0] - top2[:, 1]) > 0.1 * want32.abs().max()
    assert torch.equal(preds.float().argmax(1)[safe], want32.argmax(1)[safe])
    assert preds.argmax(1).shape == (2, 64, 96)
    net.float()
    with torch.no_grad():
        assert net(x.cuda()).dtype == torch.float32


def test_outputs_are_fresh_tensors_and_stale_backward_raises():
    from mmrseg_b200._lib import MmrError
    from mmrseg_b200.losses import DiceCrossEntropyLoss
    _, net = model_pair(2)
    xa, y = synthetic_batch(2, 2, 64, 64, seed=1)
    xb, _ = synthetic_batch(2, 2, 64, 64, seed=2)
    net.eval()
    with torch.no_grad():
        outs = [net(b.cuda()) for b in (xa, xb)]
    assert outs[0].data_ptr() != outs[1].data_ptr() and not torch.equal(outs[0], outs[1])
    with torch.no_grad():
        assert torch.equal(net(xa.cuda()), outs[0])
    net.train()
    crit = DiceCrossEntropyLoss(0.5)
    first = net(xa.cuda())
    second = net(xb.cuda())          # overwrites the plan's saved activations
    assert first.data_ptr() != second.data_ptr()
    with pytest.raises(MmrError, match="no longer the latest"):
        crit(first, y.cuda()).backward()
    crit(second, y.cuda()).backward()    # the latest forward still back-propagates


def test_eval_follows_weight_updates():
    """train -> eval -> train -> eval must score the CURRENT weights, also when they
    are rewritten by raw-pointer kernels (FusedAdam, running statistics) or load_state_dict."""
    from mmrseg_b200.losses import DiceCrossEntropyLoss
    from mmrseg_b200.optim import FusedAdam
    ref, net = model_pair(2)
    x, y = synthetic_batch(2, 2, 64, 64)
    xc, yc = x.cuda(), y.cuda()
    opt = FusedAdam(net.parameters(), lr=1e-2)
    crit = DiceCrossEntropyLoss(0.5)

    def evaluate():
        net.eval()
        with torch.no_grad():
            return net(xc)

    def oracle_eval():
        ref.load_state_dict({k: v.cpu() for k, v in net.state_dict().items()})
        ref.eval()
        with torch.no_grad():
            return ref(x)

    e0 = evaluate()
    assert rel(e0.cpu(), oracle_eval()) <= 2e-2
    for _ in range(2):
        net.train()
        for p in net.parameters():
            p.grad = None
        crit(net(xc), yc).backward()
        opt.step()
        e1 = evaluate()
        assert rel(e1, e0) > 1e-2                       # the update is visible ...
        assert rel(e1.cpu(), oracle_eval()) <= 2e-2     # ... and it is the current weights that are scored
        e0 = e1
    # load_state_dict between two eval forwards
    torch.manual_seed(99)
    from oracle.unetpp import UnetPlusPlus as OracleNet
    other = OracleNet("resnet18", None, 3, 2)
    net.load_state_dict(other.state_dict())
    e2 = evaluate()
    other.eval()
    with torch.no_grad():
        assert rel(e2.cpu(), other(x)) <= 2e-2


def test_clip_grad_norm_signature_and_parity():
    """ED/Main_MMR_SegModel.py:722: torch.nn.utils.clip_grad_norm_(self.model.parameters(), 12)."""
    from mmrseg_b200.losses import DiceCrossEntropyLoss
    from mmrseg_b200.optim import clip_grad_norm_
    _, net = model_pair(2)
    x, y = synthetic_batch(2, 2, 64, 64)
    net.train()
    DiceCrossEntropyLoss(0.5)(net(x.cuda()), y.cuda()).backward()
    grads = {k: p.grad.clone() for k, p in net.named_parameters()}
    want_norm = torch.norm(torch.stack([g.norm(2) for g in grads.values()]), 2)
    # max_norm above the norm: nothing changes
    total = clip_grad_norm_(net.parameters(), float(want_norm) * 2)
    assert total.dim() == 0 and total.is_cuda
    assert abs(float(total) - float(want_norm)) <= 1e-5 * float(want_norm)
    assert all(torch.equal(p.grad, grads[k]) for k, p in net.named_parameters())
    # max_norm below: same scaling as torch's implementation on copies of the gradients
    max_norm = float(want_norm) / 3
    copies = [torch.nn.Parameter(torch.zeros_like(g)) for g in grads.values()]
    for c, g in zip(copies, grads.values()):
        c.grad = g.clone()
    ref_total = torch.nn.utils.clip_grad_norm_(copies, max_norm)
    total = clip_grad_norm_(net.parameters(), max_norm)
    assert abs(float(total) - float(ref_total)) <= 1e-5 * float(ref_total)
    for c, (k, p) in zip(copies, net.named_parameters()):
        assert torch.allclose(p.grad, c.grad, rtol=1e-5, atol=1e-12), k
    # a subset of the parameters (one launch pair per contiguous run), and standalone tensors
    sub = [p for k, p in net.named_parameters() if k.startswith("decoder.")]
    before = {k: p.grad.clone() for k, p in net.named_parameters()}
    sub_norm = torch.norm(torch.stack([p.grad.norm(2) for p in sub]), 2)
    total = clip_grad_norm_(sub, float(sub_norm) / 2)
    assert abs(float(total) - float(sub_norm)) <= 1e-5 * float(sub_norm)
    for k, p in net.named_parameters():
        want = before[k] * (0.5 if k.startswith("decoder.") else 1.0)
        assert torch.allclose(p.grad, want, rtol=1e-5, atol=1e-12), k
    loose = [torch.nn.Parameter(torch.randn(37, device="cuda")), torch.nn.Parameter(torch.randn(5, 3, device="cuda"))]
    for p in loose:
        p.grad = torch.randn_like(p)
    copies = [torch.nn.Parameter(p.detach().clone()) for p in loose]
    for c, p in zip(copies, loose):
        c.grad = p.grad.clone()
    assert abs(float(clip_grad_norm_(loose, 0.5)) - float(torch.nn.utils.clip_grad_norm_(copies, 0.5))) <= 1e-5
    assert all(torch.allclose(p.grad, c.grad, rtol=1e-5) for p, c in zip(loose, copies))


def _train(net, opt, xc, yc, steps):
    from mmrseg_b200.losses import DiceCrossEntropyLoss
    crit = DiceCrossEntropyLoss(0.5)
    net.train()
    for _ in range(steps):
        for p in net.parameters():
            p.grad = None
        crit(net(xc), yc).backward()
        opt.step()


def test_optimizer_param_group_subsets_match_torch():
    """a group holding a subset of the flat buffer steps that subset only; two groups
    with different learning rates step every parameter once, with its own group's rate -- checked against
    torch.optim.Adam fed with the same gradients."""
    from mmrseg_b200.losses import DiceCrossEntropyLoss
    from mmrseg_b200.optim import FusedAdam
    _, net = model_pair(2)
    x, y = synthetic_batch(2, 2, 64, 64)
    xc, yc = x.cuda(), y.cuda()
    net.train()
    net._ensure_flat(xc.device)
    enc = [p for k, p in net.named_parameters() if k.startswith("encoder.")]
    rest = [p for k, p in net.named_parameters() if not k.startswith("encoder.")]
    # (1) encoder-only optimiser: decoder and head must not move
    snapshot = {k: p.detach().clone() for k, p in net.named_parameters()}
    opt = FusedAdam(enc, lr=1e-3)
    DiceCrossEntropyLoss(0.5)(net(xc), yc).backward()
    opt.step()
    for k, p in net.named_parameters():
        moved = not torch.equal(p.detach(), snapshot[k])
        assert moved == k.startswith("encoder."), k
    # (2) differential learning rates against torch.optim.Adam on clones, three steps with shared gradients
    clones = {k: torch.nn.Parameter(p.detach().clone()) for k, p in net.named_parameters()}
    c_enc = [clones[k] for k, _ in net.named_parameters() if k.startswith("encoder.")]
    c_rest = [clones[k] for k, _ in net.named_parameters() if not k.startswith("encoder.")]
    ours = FusedAdam([{"params": enc, "lr": 1e-4}, {"params": rest}], lr=1e-3, weight_decay=1e-5)
    theirs = torch.optim.Adam([{"params": c_enc, "lr": 1e-4}, {"params": c_rest}], lr=1e-3, weight_decay=1e-5)
    crit = DiceCrossEntropyLoss(0.5)
    for _ in range(3):
        for p in net.parameters():
            p.grad = None
        crit(net(xc), yc).backward()
        for k, p in net.named_parameters():
            clones[k].grad = p.grad.clone()
        ours.step()
        theirs.step()
        for k, p in net.named_parameters():
            assert torch.allclose(p.detach(), clones[k].detach(), rtol=2e-5, atol=2e-7), k
            p.data.copy_(clones[k].data)      # keep both trajectories on identical weights


def test_optimizer_state_dict_resume():
    """ADVICE round 1 (medium) / ED/Main_MMR_SegModel.py:991: optimizer.load_state_dict resumes Adam with its
    step count and moments; training A: 4 steps, training B: 2 steps, checkpoint, fresh model + optimiser,
    2 more steps -- identical parameters."""
    from mmrseg_b200.optim import FusedAdam
    x, y = synthetic_batch(2, 2, 64, 64)
    xc, yc = x.cuda(), y.cuda()
    _, a = model_pair(2)
    opt_a = FusedAdam(a.parameters(), lr=1e-3, weight_decay=1e-5)
    _train(a, opt_a, xc, yc, 4)
    _, b = model_pair(2)
    opt_b = FusedAdam(b.parameters(), lr=1e-3, weight_decay=1e-5)
    _train(b, opt_b, xc, yc, 2)
    ckpt = {"state_dict": {k: v.cpu() for k, v in b.state_dict().items()},
            "optimizer": opt_b.state_dict()}
    assert float(ckpt["optimizer"]["state"][0]["step"]) == 2.0
    _, c = model_pair(2)
    c.load_state_dict(ckpt["state_dict"])
    opt_c = FusedAdam(c.parameters(), lr=1e-3, weight_decay=1e-5)
    opt_c.load_state_dict(ckpt["optimizer"])
    _train(c, opt_c, xc, yc, 2)
    assert float(opt_c.state_dict()["state"][0]["step"]) == 4.0
    for (k, pa), (_, pc) in zip(a.named_parameters(), c.named_parameters()):
        assert torch.allclose(pa, pc, rtol=1e-4, atol=1e-6), k
    # without the optimiser state the trajectories differ (the check above is not vacuous)
    _, d = model_pair(2)
    d.load_state_dict(ckpt["state_dict"])
    _train(d, FusedAdam(d.parameters(), lr=1e-3, weight_decay=1e-5), xc, yc, 2)
    assert any(not torch.allclose(pa, pd, rtol=1e-4, atol=1e-6) for pa, pd in zip(a.parameters(), d.parameters()))


def test_amp_training_lines():
    """ED/Main_MMR_SegModel.py:696-727: autocast forward, fp32 loss, GradScaler backward / unscale_ /
    clip_grad_norm_(12) / step / update run unchanged (the kernels compute in bf16 on their own)."""
    from mmrseg_b200.losses import DiceCELoss
    from mmrseg_b200.optim import FusedAdam, clip_grad_norm_
    _, net = model_pair(10)
    images, masks = synthetic_batch(2, 10, 64, 64)
    images, masks = images.cuda(), masks.cuda()
    net.train()
    optimizer = FusedAdam(net.parameters(), lr=1e-4, weight_decay=1e-2, decoupled=True)
    scaler = torch.amp.GradScaler("cuda")
    loss_fn = DiceCELoss(softmax=True)
    before = [p.detach().clone() for p in net.parameters()]
    with torch.autocast("cuda"):
        predictions = net(images)
    masks_onehot = torch.nn.functional.one_hot(masks, 10).permute(0, 3, 1, 2)
    loss = loss_fn(predictions.float(), masks_onehot.float())
    scaler.scale(loss).backward()
    scaler.unscale_(optimizer)
    total = clip_grad_norm_(net.parameters(), 12)
    scaler.step(optimizer)
    scaler.update()
    optimizer.zero_grad()
    assert torch.isfinite(total) and float(total) > 0
    assert any(not torch.equal(b, p.detach()) for b, p in zip(before, net.parameters()))
    # the unscaled gradients equal an unscaled backward of the same step (bf16 has fp32's exponent range)
    _, net2 = model_pair(10)
    net2.train()
    loss2 = loss_fn(net2(images), masks_onehot.float())
    loss2.backward()
    assert abs(loss.item() - loss2.item()) <= 1e-6 * abs(loss2.item())


def test_native_checkpoint_on_device(tmp_path):
    """mmrseg_b200.checkpoint on a cuda model: the flat parameter buffer goes device -> file -> device, eval logits of
    the reloaded model are bit-identical, training resumes (FusedAdam state included) on the same trajectory, and
    the half-size bf16 OHWI inference checkpoint reproduces the ORIGINAL model's logits bit for bit (the kernels
    round the fp32 masters to exactly those bf16 values anyway)."""
    from mmrseg_b200 import checkpoint
    from mmrseg_b200.optim import FusedAdam
    _, net = model_pair(2)
    x, y = synthetic_batch(2, 2, 64, 64)
    xc, yc = x.cuda(), y.cuda()
    opt = FusedAdam(net.parameters(), lr=1e-3, weight_decay=1e-5)
    _train(net, opt, xc, yc, 2)
    p32, p16 = str(tmp_path / "t.mmrseg"), str(tmp_path / "i.mmrseg")
    checkpoint.save(p32, net, optimizer=opt, extra={"epoch": 2})
    checkpoint.save(p16, net, weights="bf16")
    net.eval()
    with torch.no_grad():
        want = net(xc)
    for path in (p32, p16):
        other, extra = checkpoint.load(path, device="cuda")
        other.eval()
        with torch.no_grad():
            assert torch.equal(other(xc), want), path
    # resume: two more steps from the checkpoint == two more steps of the original run
    _train(net, opt, xc, yc, 2)
    resumed, extra = checkpoint.load(p32, device="cuda")
    assert extra == {"epoch": 2}
    opt2 = FusedAdam(resumed.parameters(), lr=1.0)
    checkpoint.load(p32, resumed, optimizer=opt2)
    _train(resumed, opt2, xc, yc, 2)
    for (k, a), (_, b) in zip(net.named_parameters(), resumed.named_parameters()):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-6), k


def test_custom_ops_pass_opcheck():
    """torch.library.opcheck on the differentiable and the mutating ops: schema vs actual aliasing / mutation, fake
    implementation vs real outputs, autograd registration."""
    from torch.library import opcheck
    g = torch.Generator(device="cuda").manual_seed(0)
    logits = torch.randn((2, 3, 32, 40), device="cuda", generator=g, requires_grad=True)
    labels = torch.randint(0, 3, (2, 32, 40), device="cuda", generator=g)
    prm = (1.0, 1.0, 1e-6, 0.5, 0.5, 3, -100)
    opcheck(torch.ops.mmrseg.dice_ce_fwd, (logits, labels) + prm,
            test_utils=("test_schema", "test_faketensor", "test_autograd_registration"))
    cm = torch.zeros((2, 3, 3), device="cuda", dtype=torch.int64)
    opcheck(torch.ops.mmrseg.confusion_from_logits, (logits.detach(), labels, cm, True),
            test_utils=("test_schema", "test_faketensor"))
    p, gr, m, v = (torch.randn(1024, device="cuda", generator=g) for _ in range(4))
    opcheck(torch.ops.mmrseg.adam_step, (p, gr, m, v.abs(), 1e-3, 0.9, 0.999, 1e-8, 1e-5, 3.0, False, 1.0),
            test_utils=("test_schema", "test_faketensor"))
