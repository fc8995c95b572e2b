"""Size-independent properties at BASELINE.json's full sizes (where the CPU oracle would take minutes).

* conv adjointness on the largest U-Net++ layer (x_1_3.conv1 at batch 16 @ 256^2: one nearest-x2 source and three
  skips, 256 -> 64 channels): <conv_W(x), y> = <x, dgrad_W(y)> = <W, wgrad(x, y)>.  The three numbers come from
  three different kernels (fprop with row-phase stacking, dgrad with 128-wide N tiles, wgrad with split-K) and
  must agree to the bf16 rounding of their outputs (5e-3 relative); the nearest-x2 source's dgrad is at full
  resolution and is 2x2 sum-pooled here, as its consumer does.
* conv linearity: conv(a) + conv(b) = conv(a + b) on the same layer (bf16 outputs: 4e-3 x 8 of the RMS).
* BatchNorm statistics from the conv epilogue at full size against float64 sums of the stored tensor.
* metric conservation at 1024 x 1280 (config 5's frame size): every pixel is counted exactly once, argmax of
  one-hot logits reproduces the labels, a perfect prediction gives a diagonal matrix.
* loss at 512^2, batch 16: uniform logits give ln(C) cross-entropy and the analytic Dice value; the gradient sums to
  zero over classes at every pixel.
* one full train step at the bench configuration: finite loss that decreases over three steps on a fixed batch.
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _mk(shape, gen, scale=1.0):
    return (torch.randn(shape, generator=gen, device="cuda") * scale).to(torch.bfloat16)


def test_conv_adjointness_and_linearity_x13_full_size():
    from mmrseg_b200 import convplan
    gen = torch.Generator(device="cuda").manual_seed(6210)
    N, H, W, cout = 16, 256, 256, 64
    srcs = [(64, 2), (64, 1), (64, 1), (64, 1)]
    sources = [(_mk((N, H // up, W // up, c), gen), up) for c, up in srcs]
    cin = sum(c for c, _ in srcs)
    w = (torch.randn((cout, cin, 3, 3), generator=gen, device="cuda") / (9 * cin) ** 0.5).to(torch.bfloat16).float()
    y = _mk((N, H, W, cout), gen)
    out = torch.empty((N, H, W, cout), device="cuda", dtype=torch.bfloat16)
    stats = torch.zeros((8, 2, cout), device="cuda", dtype=torch.float64)
    fplan = convplan.build_fprop_halo(sources, w, out, stats=stats, stats_ld=cout)
    assert fplan.cfg["rph"] > 1                       # the stacked path is what runs at this size
    fplan.run()
    torch.cuda.synchronize()
    z = out.double()
    got = stats.sum(0)
    assert torch.allclose(got[0], z.sum((0, 1, 2)), rtol=1e-5, atol=1e-5 * z.abs().sum((0, 1, 2)).max().item())
    assert torch.allclose(got[1], (z * z).sum((0, 1, 2)), rtol=1e-5)
    s_f = (z * y.double()).sum().item()
    # dgrad: one gradient tensor per source at the conv's resolution
    grads = [torch.empty((N, H, W, c), device="cuda", dtype=torch.bfloat16) for c, _ in srcs]
    dplan = convplan.build_dgrad_halo(y, w, grads)
    dplan.run()
    torch.cuda.synchronize()
    s_d = 0.0
    for (t, up), g in zip(sources, grads):
        gd = g.double()
        if up == 2:
            gd = gd.view(N, H // 2, 2, W // 2, 2, -1).sum((2, 4))
        s_d += (gd * t.double()).sum().item()
    dw = torch.empty((cout, cin, 3, 3), device="cuda")
    wplan = convplan.build_wgrad_halo(y, sources, dw)
    wplan.run()
    torch.cuda.synchronize()
    s_w = (dw.double() * w.double()).sum().item()
    scale = (z.abs() * y.double().abs()).sum().item()
    assert abs(s_f - s_d) <= 5e-3 * scale / 50 and abs(s_f - s_w) <= 5e-3 * scale / 50, (s_f, s_d, s_w, scale)
    assert abs(s_f - s_d) <= 5e-3 * abs(s_f) + 1e-6 * scale and abs(s_f - s_w) <= 5e-3 * abs(s_f) + 1e-6 * scale
    # linearity
    sources_b = [(_mk(t.shape, gen), up) for t, up in sources]
    sources_ab = [((a.float() + b.float()).to(torch.bfloat16), up) for (a, up), (b, _) in zip(sources, sources_b)]
    out_b, out_ab = torch.empty_like(out), torch.empty_like(out)
    convplan.build_fprop_halo(sources_b, w, out_b).run()
    convplan.build_fprop_halo(sources_ab, w, out_ab).run()
    torch.cuda.synchronize()
    # a + b is itself rounded to bf16 before the conv: compare against the conv of the exact sum's rounding error
    diff = (out.float() + out_b.float() - out_ab.float())
    rms = out_ab.float().pow(2).mean().sqrt().item()
    assert diff.abs().max().item() <= 6e-2 * rms * 8, (diff.abs().max().item(), rms)
    assert diff.pow(2).mean().sqrt().item() <= 1.5e-2 * rms


def test_metric_conservation_at_endoscopic_resolution():
    from mmrseg_b200.metrics import confusion_matrix, dice_per_image
    g = torch.Generator(device="cuda").manual_seed(5)
    N, C, H, W = 4, 10, 1024, 1280
    labels = torch.randint(0, C, (N, H, W), generator=g, device="cuda")
    logits = torch.randn((N, C, H, W), generator=g, device="cuda")
    cm, pred = confusion_matrix(logits, labels, return_pred=True)
    assert int(cm.sum()) == N * H * W and torch.equal(cm.sum((1, 2)), torch.full((N,), H * W, device="cuda"))
    assert torch.equal(cm.sum(2), torch.stack([torch.bincount(labels[i].flatten(), minlength=C) for i in range(N)]))
    assert torch.equal(cm.sum(1), torch.stack([torch.bincount(pred[i].flatten(), minlength=C) for i in range(N)]))
    onehot = F.one_hot(labels, C).permute(0, 3, 1, 2).float().contiguous()
    cm2, pred2 = confusion_matrix(onehot, labels, return_pred=True)
    assert torch.equal(pred2, labels)
    assert torch.equal(cm2, torch.diag_embed(cm2.diagonal(dim1=1, dim2=2)))
    assert torch.equal(dice_per_image(pred2, labels, C), torch.ones(N, dtype=torch.float64, device="cuda"))


def test_loss_closed_forms_at_full_size():
    from mmrseg_b200.losses import DiceCrossEntropyLoss
    N, C, H, W = 16, 2, 512, 512
    g = torch.Generator(device="cuda").manual_seed(7)
    labels = torch.randint(0, C, (N, H, W), generator=g, device="cuda")
    logits = torch.zeros((N, C, H, W), device="cuda", requires_grad=True)
    loss = DiceCrossEntropyLoss(0.5)(logits, labels)
    loss.backward()
    # uniform softmax p = 1/C: CE = ln C; dice term = mean_{n,c} 1 - (2 I + 1) / (Card + 1) with the kornia +1e-6 one-hot
    hw = H * W
    cnt = torch.stack([(labels == c).sum((1, 2)) for c in range(C)], 1).double()
    inter = (cnt + 1e-6 * hw) / C
    card = hw / C + cnt + 1e-6 * hw
    dice = (1 - (2 * inter + 1.0) / (card + 1.0)).mean().item()
    want = 0.5 * dice + 0.5 * math.log(C)
    assert abs(loss.item() - want) <= 1e-5 * want, (loss.item(), want)
    assert logits.grad.sum(1).abs().max().item() <= 1e-9        # softmax Jacobian: zero over classes at every pixel


def test_train_steps_at_bench_configuration():
    import bench
    from mmrseg_b200.losses import DiceCrossEntropyLoss
    from mmrseg_b200.models import UnetPlusPlus
    from mmrseg_b200.optim import FusedAdam
    torch.manual_seed(6210)
    model = UnetPlusPlus("resnet18", classes=2).cuda().train()
    crit = DiceCrossEntropyLoss(0.5)
    opt = FusedAdam(model.parameters(), lr=1e-3, weight_decay=1e-5)
    cfg = bench.resolve("c2", 1)
    x, y = bench.synthetic(cfg, cfg["batch"])
    x, y = x.cuda(), y.cuda()
    losses = []
    for _ in range(4):          # eager, capture, two graph replays
        for p in model.parameters():
            p.grad = None
        loss = crit(model(x), y)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert all(math.isfinite(v) for v in losses) and losses[-1] < losses[0], losses
    eng = list(model._engines.values())[0]
    # every BatchNorm input the plan keeps has the batch statistics the plan recorded
    for name in ("decoder.blocks.x_1_3.conv1.0", "encoder.layer2.1.conv2", "encoder.conv1"):
        u = next(u for u in eng.units if u.get("op", {}).get("conv") == name)
        z = u["z"].double().reshape(-1, u["cout"])
        assert torch.allclose(z.mean(0).float(), u["mean"], atol=2e-3, rtol=2e-3), name
        assert torch.allclose((z.var(0, unbiased=False) + 1e-5).rsqrt().float(), u["invstd"], rtol=2e-3), name
