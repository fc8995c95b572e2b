"""Whole-model parity at BASELINE sizes ("next" #1).

1. Teacher-forced layer-local parity (tests/teacher.py explains why this, and not a tighter end-to-end number,
   is what separates arithmetic from bugs): every unit of the plan recomputed in fp32 PyTorch on the CPU from
   the tensors the engine actually consumed.  Tolerances (relative Frobenius): bf16 outputs 5e-4 (one-ulp flips
   of the summation order only; measured worst 2.2e-4), fp32 outputs 1e-4 (measured: weight gradients <= 3.2e-5
   with fp32 accumulation over 5e5 pixels in another order, statistics / logits / BatchNorm gradients ~1e-7),
   max-pool values exact (profiles/r02_parity_teacher.txt).  Run on BASELINE's own shapes:
   U-Net++-R18 @ 512x512 (configs 2 / 4, batch 2 and deep supervision), ResNetUNet-34 @ 512x512 (config 3),
   plus resnet34 / UNet small cases.
2. End to end against the plain fp32 oracle at 512 x 512 (no masks injected): logits <= 1e-1, loss <= 2e-3,
   and the statement that makes the loose gradient numbers meaningful -- the fp32 oracle with bf16 rounding at the
   engine's storage points (oracle/bf16_points.py) is as far from the plain fp32 oracle as the engine is.
3. Config 5's shape: eval forward of one 1024 x 1280 frame, 10 classes, against the fp32 oracle (logits <= 2e-2)
   and bit-exact argmax / confusion counts on the engine's own logits.
"""
import numpy as np
import pytest
import torch

from tests import teacher
from tests.helpers import model_pair, rel, synthetic_batch

pytestmark = pytest.mark.gpu

TOL = {"bf16": 5e-4, "f32": 1e-4, "exact": 0.0, "zero": 1e-3}


def _teacher_forced(net, x, y, monkeypatch, report=None):
    from mmrseg_b200.losses import DiceCrossEntropyLoss
    monkeypatch.setenv("MMR_NO_ARENA_REUSE", "1")      # every backward temporary keeps its own memory
    net.train()
    P = teacher.snapshot(net)
    out = net(x.cuda())
    outs = out if isinstance(out, list) else [out]
    crit = DiceCrossEntropyLoss(0.5)
    loss = sum(crit(o, y.cuda()) for o in outs) / len(outs)
    loss.backward()
    torch.cuda.synchronize()
    eng = [e for k, e in net._engines.items() if k[3]][0]
    rows = teacher.check_forward(eng, P, x) + teacher.check_backward(eng, P, teacher.grads_of(net), x)
    assert len(rows) > 3 * len([u for u in eng.units if u["kind"] in ("conv", "stem", "head")])
    worst = teacher.summarize(rows)
    if report is not None:
        report(rows, worst)
    for unit, what, kind, err in rows:
        assert err <= TOL[kind], (unit, what, kind, err, worst)
    return worst


@pytest.mark.parametrize("encoder,classes,n,h,w,ds", [
    ("resnet18", 2, 2, 512, 512, False),      # BASELINE config 2's shape (batch 2 of 16)
    ("resnet18", 2, 1, 512, 512, True),       # config 4: deep supervision
    ("resnet34", 10, 2, 128, 160, False),
    ("resnet18", 10, 3, 64, 96, False),
])
def test_teacher_forced_unetpp(encoder, classes, n, h, w, ds, monkeypatch):
    if ds:
        from oracle.unetpp import DeepSupervisionUnetPlusPlus
        from mmrseg_b200.models import UnetPlusPlus
        torch.manual_seed(6210)
        ref = DeepSupervisionUnetPlusPlus(encoder, None, 3, classes)
        net = UnetPlusPlus(encoder, classes=classes, deep_supervision=True)
        net.load_state_dict(ref.state_dict(), strict=True)
        net = net.cuda()
    else:
        _, net = model_pair(classes, encoder)
    x, y = synthetic_batch(n, classes, h, w)
    worst = _teacher_forced(net, x, y, monkeypatch)
    print("teacher-forced U-Net++ %s %s: worst %s" % (encoder, (n, h, w), worst))


def test_teacher_forced_resnet_unet34_config3_shape(monkeypatch):
    from tests.test_resnet_unet_gpu import _pair
    _, net = _pair(10, 34)
    x, y = synthetic_batch(1, 10, 512, 512)
    worst = _teacher_forced(net.cuda(), x, y, monkeypatch)
    print("teacher-forced ResNetUNet-34 (1, 512, 512): worst %s" % (worst,))


def test_teacher_forced_unet(monkeypatch):
    from tests.test_unet_gpu import _pair
    _, net = _pair(10)
    x, y = synthetic_batch(2, 10, 128, 96)
    worst = _teacher_forced(net.cuda(), x, y, monkeypatch)
    print("teacher-forced UNet (2, 128, 96): worst %s" % (worst,))


def test_teacher_forced_eval_mode():
    ref, net = model_pair(10)
    x, _ = synthetic_batch(1, 10, 256, 320)
    net.eval()
    P = teacher.snapshot(net)
    with torch.no_grad():
        net(x.cuda())
    eng = [e for k, e in net._engines.items() if not k[3]][0]
    rows = teacher.check_forward(eng, P, x, training=False)
    for unit, what, kind, err in rows:
        assert err <= TOL[kind], (unit, what, kind, err)


def test_end_to_end_fp32_oracle_at_512():
    """Config 2's resolution, batch 2 (the CPU oracle needs ~1 s): the engine against the plain fp32 oracle with
    nothing injected, and the bf16-rounding-points oracle against the same fp32 oracle as the yardstick."""
    from oracle import bf16_points
    from oracle.losses import mixed_loss
    from mmrseg_b200.losses import DiceCrossEntropyLoss
    ref, net = model_pair(2)
    x, y = synthetic_batch(2, 2, 512, 512)
    ref.train()
    net.train()
    got = net(x.cuda())
    loss = DiceCrossEntropyLoss(0.5)(got, y.cuda())
    loss.backward()
    eng_g = teacher.grads_of(net)
    sd = {k: v.clone() for k, v in ref.state_dict().items()}
    want = ref(x)
    loss_ref = mixed_loss(want, y, 0.5)
    loss_ref.backward()
    ref_g = {k: p.grad.clone() for k, p in ref.named_parameters()}
    ref.load_state_dict(sd)
    ref.zero_grad(set_to_none=True)
    mid = bf16_points.unetpp_forward(ref, x)
    mixed_loss(mid, y, 0.5).backward()
    mid_g = {k: p.grad.clone() for k, p in ref.named_parameters()}
    e_logits, m_logits = rel(got.detach().cpu(), want.detach()), rel(mid.detach(), want.detach())
    assert e_logits <= 1e-1, e_logits
    assert abs(loss.item() - loss_ref.item()) <= 2e-3 * abs(loss_ref.item())
    # bf16 rounding alone (a CPU-only statement about the arithmetic) costs as much as the engine's total error
    assert e_logits <= 1.5 * m_logits + 1e-3, (e_logits, m_logits)
    e_med = sorted(rel(eng_g[k], ref_g[k]) for k in ref_g)[len(ref_g) // 2]
    m_med = sorted(rel(mid_g[k], ref_g[k]) for k in ref_g)[len(ref_g) // 2]
    assert e_med <= 1.5 * m_med + 1e-2, (e_med, m_med)
    # the head and the last decoder block see few rounding points: their gradients are tight even end to end
    for k in ("segmentation_head.0.weight", "decoder.blocks.x_0_4.conv2.0.weight"):
        assert rel(eng_g[k], ref_g[k]) <= 6e-2, (k, rel(eng_g[k], ref_g[k]))
    print("512x512 end to end vs fp32 oracle: logits %.3e (bf16-points oracle %.3e), median grad %.3e (%.3e)" % (
        e_logits, m_logits, e_med, m_med))


def test_config5_shape_eval_and_bit_exact_confusion():
    """One 1024 x 1280 frame, 10 classes, eval mode (BN folded): logits against the fp32 oracle, and the metric
    path (argmax, confusion matrix, Evaluate tp / fp / fn) bit-exact on the engine's own logits."""
    from oracle import metrics as OM
    from mmrseg_b200.metrics import Evaluate, confusion_matrix
    ref, net = model_pair(10)
    x, y = synthetic_batch(1, 10, 1024, 1280)
    ref.eval()
    net.eval()
    with torch.no_grad():
        want = ref(x)
        logits = net(x.cuda())
    assert rel(logits.cpu(), want) <= 2e-2, rel(logits.cpu(), want)
    cm, pred = confusion_matrix(logits, y.cuda(), return_pred=True)
    lg = logits.cpu().numpy()
    want_pred = OM.argmax_first(lg)
    assert np.array_equal(pred.cpu().numpy(), want_pred)
    want_cm = OM.confusion_matrix(want_pred, y.numpy(), 10)      # int64 [1, 10, 10]
    assert np.array_equal(cm.cpu().numpy(), want_cm)
    ev = Evaluate({i: i for i in range(10)}, use_gpu=True)

    class A:
        dataset = "sarrarp50"
    onehot = torch.nn.functional.one_hot(y, 10).permute(0, 3, 1, 2).cuda()
    ev.addBatch(logits, onehot, A())
    full = want_cm.sum(0)
    tp = np.diag(full).astype(np.float64)
    assert np.array_equal(ev.tp.numpy(), tp)
    want_iou = tp / (full.sum(0) + full.sum(1) - tp + 1e-15)
    assert np.abs(ev.getIoU().numpy() - want_iou).max() <= 1e-12


@pytest.mark.parametrize("classes,shape,u8_labels", [(10, (2, 96, 160), False), (2, (3, 64, 64), False), (16, (1, 64, 96), False)])
def test_fused_head_metric_is_bit_exact(classes, shape, u8_labels):
    """argmax + confusion matrix in the head's epilogue (model.segment) against the two-kernel path
    (forward -> logits in HBM -> mmr_confusion_from_logits) and the int64 numpy oracle on the same logits:
    predictions and counts identical, including labels outside [0, C) (skipped) and an accumulating Evaluate."""
    from oracle import metrics as OM
    from mmrseg_b200.metrics import Evaluate, confusion_matrix
    _, net = model_pair(classes)
    n, h, w = shape
    x, y = synthetic_batch(n, classes, h, w)
    y[0, :3, :5] = classes + 3          # out of range: not counted
    y[0, 3, :5] = -1
    net.eval()
    with torch.no_grad():
        logits = net(x.cuda())
    cm2, pred2 = confusion_matrix(logits, y.cuda(), return_pred=True)
    pred, cm = net.segment(x.cuda(), y.cuda())
    assert pred.dtype == torch.uint8 and cm.dtype == torch.int64
    assert torch.equal(pred.long(), pred2) and torch.equal(cm, cm2)
    want_pred = OM.argmax_first(logits.cpu().numpy())
    assert np.array_equal(pred.cpu().numpy(), want_pred)
    assert np.array_equal(cm.cpu().numpy(), OM.confusion_matrix(want_pred, y.numpy(), classes))
    assert int(cm.sum()) == n * h * w - 20
    # a second call starts from zero again; without labels nothing is counted
    pred_b, cm_b = net.segment(x.cuda(), y.cuda())
    assert torch.equal(cm_b, cm) and torch.equal(pred_b, pred)
    pred_c, none = net.segment(x.cuda())
    assert none is None and torch.equal(pred_c, pred)
    ev = Evaluate({i: i for i in range(classes)}, use_gpu=True)
    for _ in range(2):
        ev.addBatchFromModel(net, x.cuda(), y.cuda())
    assert torch.equal(ev.confusion(), 2 * cm.sum(0))
