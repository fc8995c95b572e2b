"""Loss and metric kernels through the reference-facing API against the oracle and the
reference's golden vectors.  Loss: |delta| <= 1e-5 relative, dlogits <= 1e-6 abs + 1e-4 rel (fp32).
Metric: argmax masks and confusion counts bit-exact (int64)."""
import os
import types

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _case(n, c, h, w, seed, scale=2.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn((n, c, h, w), generator=g) * scale, torch.randint(0, c, (n, h, w), generator=g)


@pytest.mark.parametrize("n,c,h,w", [(2, 2, 32, 32), (3, 10, 24, 40), (1, 5, 7, 9), (16, 2, 64, 64), (2, 21, 16, 16)])
def test_dice_loss_matches_oracle(n, c, h, w):
    from oracle import losses as O
    from mmrseg_b200 import losses as L
    z, t = _case(n, c, h, w, 1)
    zr = z.clone().requires_grad_(True)
    want = O.dice_loss(zr, t)
    want.backward()
    zc = z.cuda().requires_grad_(True)
    got = L.DiceLoss()(zc, t.cuda())
    got.backward()
    assert abs(got.item() - want.item()) <= 1e-5 * abs(want.item())
    assert torch.allclose(zc.grad.cpu(), zr.grad, rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize("w_dice", [0.5, 0.3, -1])
@pytest.mark.parametrize("c", [2, 10])
def test_mixed_loss_matches_oracle(w_dice, c):
    from oracle import losses as O
    from mmrseg_b200 import losses as L
    z, t = _case(4, c, 32, 48, 2)
    zr = z.clone().requires_grad_(True)
    want = O.mixed_loss(zr, t, w_dice)
    (want * 3.0).backward()
    zc = z.cuda().requires_grad_(True)
    got = L.DiceCrossEntropyLoss(w_dice)(zc, t.cuda())
    (got * 3.0).backward()          # upstream gradient reaches the kernel as a device scalar
    assert abs(got.item() - want.item()) <= 1e-5 * abs(want.item())
    assert torch.allclose(zc.grad.cpu(), zr.grad, rtol=1e-4, atol=1e-7)


def test_dice_ignore_index_slice_and_ce_ignore():
    from oracle import losses as O
    from mmrseg_b200 import losses as L
    z, t = _case(2, 6, 16, 16, 3)
    want = O.dice_loss(z, t, 1.0, ignore_index=4)
    got = L.dice_loss(z.cuda(), t.cuda(), 1.0, ignore_index=4)
    assert abs(got.item() - want.item()) <= 1e-5 * abs(want.item())
    t2 = t.clone()
    t2[0, :4] = 21                                            # ignore_index=21 ("synapse", SU/ModelTraining.py:358)
    z21, _ = _case(2, 22, 16, 16, 4)
    want = F.cross_entropy(z21, t2, ignore_index=21)
    got = L.DiceCrossEntropyLoss(-1, ce_ignore_index=21)(z21.cuda(), t2.cuda())
    assert abs(got.item() - want.item()) <= 1e-5 * abs(want.item())


def test_monai_dice_ce_matches_oracle():
    from oracle import losses as O
    from mmrseg_b200 import losses as L
    z, t = _case(3, 10, 32, 32, 5)
    oh = F.one_hot(t, 10).permute(0, 3, 1, 2).float()
    zr = z.clone().requires_grad_(True)
    want = O.monai_dice_ce(zr, oh)
    want.backward()
    zc = z.cuda().requires_grad_(True)
    got = L.DiceCELoss(softmax=True)(zc, oh.cuda())
    got.backward()
    assert abs(got.item() - want.item()) <= 1e-5 * abs(want.item())
    assert torch.allclose(zc.grad.cpu(), zr.grad, rtol=1e-4, atol=1e-7)


def test_loss_error_behaviour():
    from mmrseg_b200 import losses as L
    from mmrseg_b200._lib import MmrError
    t = torch.zeros((1, 4, 4), dtype=torch.long, device="cuda")
    with pytest.raises(TypeError):
        L.dice_loss([1.0], t)
    with pytest.raises(ValueError):
        L.dice_loss(torch.zeros(1, 2, 4, device="cuda"), t)
    with pytest.raises(ValueError):
        L.dice_loss(torch.zeros(1, 2, 5, 5, device="cuda"), t)
    with pytest.raises(ValueError):
        L.dice_loss(torch.zeros(1, 2, 4, 4), t)                # device mismatch
    with pytest.raises(MmrError):
        L.dice_loss(torch.zeros(1, 2, 4, 4), t.cpu())          # CPU tensors: no fallback


@pytest.mark.parametrize("name", ["c2", "c10", "c5", "ties"])
def test_evaluate_matches_reference_golden(name):
    from mmrseg_b200.metrics import Evaluate, dice
    g = np.load(os.path.join(GOLD, "metrics_reference.npz"))
    logits = torch.from_numpy(g[name + "_logits"]).cuda()
    labels = torch.from_numpy(g[name + "_labels"]).cuda()
    c = logits.shape[1]
    ev = Evaluate({i: None for i in range(c)}, use_gpu=True)
    onehot = F.one_hot(labels, c).permute(0, 3, 1, 2)
    args = types.SimpleNamespace(dataset="sarrarp50")
    ev.addBatch(logits, onehot, args)
    ev.addBatch(logits.flip(0), onehot.flip(0), args)
    assert np.array_equal(ev.tp.numpy(), g[name + "_tp"])
    assert np.array_equal(ev.fp.numpy(), g[name + "_fp"])
    assert np.array_equal(ev.fn.numpy(), g[name + "_fn"])
    assert np.allclose(ev.getIoU().numpy(), g[name + "_iou"], rtol=0, atol=1e-12)
    p, r, f1 = ev.getPRF1()
    assert np.allclose(f1.numpy(), g[name + "_f1"], rtol=0, atol=1e-12)
    pred = torch.argmax(logits, 1)
    d = [dice(F.one_hot(pred[i], c).permute(2, 0, 1), onehot[i]) for i in range(len(logits))]
    d += [dice(torch.zeros(4, 4, device="cuda"), torch.zeros(4, 4, device="cuda")),
          dice(torch.zeros(4, 4, device="cuda"), torch.zeros(4, 4, device="cuda"), 0.5)]
    assert np.allclose(d, g[name + "_dice"], atol=1e-15)
    ev.reset()
    assert ev.confusion() is None


def test_confusion_bit_exact_vs_oracle_including_ties_and_nan():
    from oracle import metrics as O
    from mmrseg_b200.metrics import confusion_matrix
    g = torch.Generator().manual_seed(9)
    n, c, h, w = 3, 10, 96, 160
    logits = torch.randn((n, c, h, w), generator=g).to(torch.bfloat16).float()   # bf16 grid: many ties
    logits[0, :, :8] = 0.0                                                       # exact all-way ties
    labels = torch.randint(0, c, (n, h, w), generator=g)
    cm, pred = confusion_matrix(logits.cuda(), labels.cuda(), return_pred=True)
    assert torch.equal(pred.cpu(), torch.argmax(logits, 1))
    want = O.confusion_matrix(O.argmax_first(logits.numpy()), labels.numpy(), c)
    assert np.array_equal(cm.cpu().numpy(), want)
    assert int(cm.sum()) == n * h * w                                            # checksum of checksums


def test_confusion_large_counts_exceed_float32():
    """Above 2^24 pixels per class the reference's float32 counting is inexact;
    the int64 kernel stays exact: closed-form check on a constant prediction."""
    from mmrseg_b200.metrics import confusion_matrix
    n, c, h, w = 2, 3, 4096, 4096
    logits = torch.zeros((n, c, h, w), device="cuda")
    logits[:, 1] = 1.0
    labels = torch.ones((n, h, w), dtype=torch.long, device="cuda")
    labels[:, :, :1] = 2
    cm = confusion_matrix(logits, labels).sum(0).cpu()
    assert int(cm[1, 1]) == n * h * (w - 1) and int(cm[1, 1]) > 2 ** 24
    assert int(cm[2, 1]) == n * h and int(cm.sum()) == n * h * w


def test_get_stats_iou_score_match_oracle():
    from oracle import metrics as O
    from mmrseg_b200.metrics import get_stats, iou_score
    g = torch.Generator().manual_seed(11)
    pred = torch.randint(0, 10, (4, 64, 80), generator=g)
    mask = torch.randint(0, 10, (4, 64, 80), generator=g)
    tp, fp, fn, tn = get_stats(pred.cuda(), mask.cuda(), mode="multiclass", num_classes=10)
    wtp, wfp, wfn, wtn = O.get_stats(pred.numpy(), mask.numpy(), 10)
    for a, b in ((tp, wtp), (fp, wfp), (fn, wfn), (tn, wtn)):
        assert a.dtype == torch.int64 and np.array_equal(a.cpu().numpy(), b)
    macro = iou_score(tp, fp, fn, tn, reduction="macro")
    assert abs(float(macro) - float(O.iou_score(wtp, wfp, wfn, wtn, "macro"))) <= 1e-6
    # inference call: background shifted to -1 and ignored (ED/Main_MMR_SegModel.py:1323-1325)
    tp, fp, fn, tn = get_stats(pred.cuda() - 1, mask.cuda() - 1, mode="multiclass", num_classes=9, ignore_index=-1)
    wtp, wfp, wfn, wtn = O.get_stats(pred.numpy() - 1, mask.numpy() - 1, 9, ignore_index=-1)
    for a, b in ((tp, wtp), (fp, wfp), (fn, wfn), (tn, wtn)):
        assert np.array_equal(a.cpu().numpy(), b)
    assert np.allclose(iou_score(tp, fp, fn, tn).cpu().numpy(), O.iou_score(wtp, wfp, wfn, wtn), atol=1e-6)


def test_adam_matches_torch():
    from mmrseg_b200.optim import FusedAdam
    g = torch.Generator().manual_seed(12)
    for decoupled in (False, True):
        flat = torch.randn(10007, generator=g).cuda()
        p1 = [torch.nn.Parameter(flat[:5000].clone()), torch.nn.Parameter(flat[5000:].clone())]
        holder = flat.clone()
        p2 = [torch.nn.Parameter(holder[:5000]), torch.nn.Parameter(holder[5000:])]
        ref = (torch.optim.AdamW if decoupled else torch.optim.Adam)(p1, lr=1e-3, weight_decay=1e-2)
        opt = FusedAdam(p2, lr=1e-3, weight_decay=1e-2, decoupled=decoupled)
        for step in range(3):
            gr = torch.randn(10007, generator=g).cuda()
            gh = gr.clone()
            p1[0].grad, p1[1].grad = gr[:5000].clone(), gr[5000:].clone()
            p2[0].grad, p2[1].grad = gh[:5000], gh[5000:]
            ref.step()
            opt.step()
        got = torch.cat([p.data for p in p2])
        want = torch.cat([p.data for p in p1])
        assert torch.allclose(got, want, rtol=1e-5, atol=1e-7)


def test_sgd_matches_torch():
    """FusedSGD against torch.optim.SGD (the reference's optim.SGD(..., momentum=0.9), SU/ModelTraining.py:372,381):
    one flat launch, per-group learning rates, with and without momentum / weight decay."""
    from mmrseg_b200.optim import FusedSGD
    g = torch.Generator().manual_seed(13)
    for momentum, wd, groups in ((0.9, 0.0, False), (0.9, 1e-2, True), (0.0, 1e-3, False)):
        flat = torch.randn(10007, generator=g).cuda()
        p1 = [torch.nn.Parameter(flat[:5000].clone()), torch.nn.Parameter(flat[5000:].clone())]
        holder = flat.clone()
        p2 = [torch.nn.Parameter(holder[:5000]), torch.nn.Parameter(holder[5000:])]
        if groups:   # differential learning rates (SU/ModelTraining.py:375-383)
            ref = torch.optim.SGD([{"params": [p1[0]], "lr": 1e-2}, {"params": [p1[1]]}], lr=1e-3, momentum=momentum,
                                  weight_decay=wd)
            opt = FusedSGD([{"params": [p2[0]], "lr": 1e-2}, {"params": [p2[1]]}], lr=1e-3, momentum=momentum,
                           weight_decay=wd)
        else:
            ref = torch.optim.SGD(p1, lr=1e-2, momentum=momentum, weight_decay=wd)
            opt = FusedSGD(p2, lr=1e-2, momentum=momentum, weight_decay=wd)
        for step in range(4):
            gr = torch.randn(10007, generator=g).cuda()
            gh = gr.clone()
            p1[0].grad, p1[1].grad = gr[:5000].clone(), gr[5000:].clone()
            p2[0].grad, p2[1].grad = gh[:5000], gh[5000:]
            ref.step()
            opt.step()
        got = torch.cat([p.data for p in p2])
        want = torch.cat([p.data for p in p1])
        assert torch.allclose(got, want, rtol=1e-5, atol=1e-7), (momentum, wd, groups)


def test_dice_per_image_matches_reference_dice_on_one_hots():
    """metrics.dice_per_image against oracle.metrics.dice (pinned to the reference's utils.dice by
    tests/golden/metrics_reference.npz) applied to the one-hot volumes, as SU/ModelTraining.py:629-634 does."""
    import numpy as np
    from oracle import metrics as OM
    from mmrseg_b200.metrics import dice_per_image
    g = torch.Generator().manual_seed(12)
    n, c, h, w = 3, 5, 20, 28
    pred = torch.randint(0, c, (n, h, w), generator=g)
    gt = torch.randint(0, c, (n, h, w), generator=g)
    gt[2] = pred[2]                                   # a perfect image
    got = dice_per_image(pred.cuda(), gt.cuda(), c).cpu().numpy()
    for i in range(n):
        a = torch.nn.functional.one_hot(pred[i], c).permute(2, 0, 1).numpy()
        b = torch.nn.functional.one_hot(gt[i], c).permute(2, 0, 1).numpy()
        assert abs(got[i] - OM.dice(a, b)) <= 1e-12, i
    assert got[2] == 1.0


def _blobs(n, classes, h, w, seed):
    """Labels with spatial structure (background + random ellipses, SURVEY 8d variant B) and a perturbed copy."""
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    maps = []
    for _ in range(2):
        m = torch.zeros((n, h, w), dtype=torch.int64)
        for i in range(n):
            for _ in range(6):
                c = int(torch.randint(1, classes, (1,), generator=g))
                cy, cx = float(torch.rand(1, generator=g)) * h, float(torch.rand(1, generator=g)) * w
                ry, rx = 2 + float(torch.rand(1, generator=g)) * h / 5, 2 + float(torch.rand(1, generator=g)) * w / 5
                m[i][((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1] = c
        maps.append(m)
    return maps


@pytest.mark.parametrize("n,classes,h,w", [(3, 4, 40, 56), (2, 10, 96, 64), (1, 2, 33, 130)])
def test_hausdorff_matches_skimage_restatement(n, classes, h, w):
    """Every image and class against the restated skimage.metrics.hausdorff_distance (two cKDTree queries): the
    device computes the exact squared distance in integers, so the float64 results are identical; covers classes
    absent from one map (inf), from both (0) and the reference's cap of inf at 1000 (SU/ModelTraining.py:646)."""
    from oracle import metrics as OM
    from mmrseg_b200.metrics import detailed_metrics, hausdorff_distance
    label, pred = _blobs(n, classes, h, w, seed=n * 100 + classes)
    pred[0][pred[0] == 1] = 0                 # class 1 missing from the prediction of image 0 -> inf (if labelled)
    if classes > 3:
        label[:, :, :][label == 3] = 0        # class 3 labelled nowhere ...
        pred[:, :, :][pred == 3] = 0          # ... and predicted nowhere -> 0
    for p_dev in (pred.cuda(), pred.to(torch.uint8).cuda()):
        got = hausdorff_distance(p_dev, label.cuda(), classes).cpu().numpy()
        for i in range(n):
            for c in range(classes):
                want = OM.hausdorff_distance(pred[i].numpy() == c, label[i].numpy() == c)
                assert got[i, c] == want, (i, c, got[i, c], want)
    dice, hd = detailed_metrics(pred.cuda(), label.cuda(), classes)
    # the reference's accumulation: dice of the C x H x W one-hot volumes, Hausdorff with inf -> 1000
    oh = lambda m: torch.nn.functional.one_hot(m, classes).permute(2, 0, 1).numpy()
    total_dice = sum(OM.dice(oh(pred[i]), oh(label[i])) for i in range(n))
    total_haus = 0.0
    for i in range(n):
        for c in range(classes):
            v = OM.hausdorff_distance(oh(pred[i])[c], oh(label[i])[c])
            total_haus += 1000 if v == np.inf else v
    assert abs(float(dice.sum()) - total_dice) <= 1e-12 * max(1.0, total_dice)
    assert abs(float(hd.sum()) - total_haus) <= 1e-9 * max(1.0, total_haus)
