"""The oracle against the reference's own pins: golden vectors produced by running
the reference's code (oracle/make_golden.py), and the structural pins the reference publishes."""
import os
import types

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", ["c2", "c10", "c5", "ties"])
def test_evaluate_and_dice_match_reference_golden(name):
    from oracle import metrics
    g = np.load(os.path.join(GOLD, "metrics_reference.npz"))
    logits, labels = g[name + "_logits"], g[name + "_labels"]
    c = logits.shape[1]
    ev = metrics.Evaluate({i: None for i in range(c)})
    onehot = np.eye(c, dtype=np.int64)[labels].transpose(0, 3, 1, 2)
    args = types.SimpleNamespace(dataset="sarrarp50")
    ev.addBatch(logits, onehot, args)
    ev.addBatch(logits[::-1], onehot[::-1], args)
    assert np.array_equal(ev.tp, g[name + "_tp"])
    assert np.array_equal(ev.fp, g[name + "_fp"])
    assert np.array_equal(ev.fn, g[name + "_fn"])
    assert np.allclose(ev.getIoU(), g[name + "_iou"], rtol=0, atol=1e-12)
    p, r, f1 = ev.getPRF1()
    assert np.allclose(p, g[name + "_p"], atol=1e-12) and np.allclose(r, g[name + "_r"], atol=1e-12)
    assert np.allclose(f1, g[name + "_f1"], atol=1e-12)
    pred = metrics.argmax_first(logits)
    d = [metrics.dice(np.eye(c)[pred[i]].transpose(2, 0, 1), onehot[i]) for i in range(len(logits))]
    d += [metrics.dice(np.zeros((4, 4)), np.zeros((4, 4))), metrics.dice(np.zeros((4, 4)), np.zeros((4, 4)), 0.5)]
    assert np.allclose(d, g[name + "_dice"], atol=1e-15)


def test_confusion_matrix_int64_properties():
    from oracle import metrics
    rng = np.random.default_rng(0)
    pred = rng.integers(0, 7, (3, 50, 40))
    lab = rng.integers(0, 7, (3, 50, 40))
    cm = metrics.confusion_matrix(pred, lab, 7)
    assert cm.dtype == np.int64 and cm.sum() == pred.size
    assert np.array_equal(cm.sum(2), np.stack([np.bincount(l.ravel(), minlength=7) for l in lab]))
    tp, fp, fn, tn = metrics.get_stats(pred, lab, 7)
    assert np.array_equal(tp, np.diagonal(cm, axis1=1, axis2=2))
    assert np.array_equal(tp + fp + fn + tn, np.full((3, 7), 2000))
    # ignore_index = -1 after the reference's "preds-1, masks-1" shift (ED/Main_MMR_SegModel.py:1323)
    tp2, fp2, fn2, tn2 = metrics.get_stats(pred - 1, lab - 1, 6, ignore_index=-1)
    assert np.array_equal(tp2, tp[:, 1:])
    iou = metrics.iou_score(tp2, fp2, fn2, tn2)
    assert iou.shape == (3, 6) and np.all((iou >= 0) & (iou <= 1))
    z = np.zeros((1, 3), dtype=np.int64)
    assert np.array_equal(metrics.iou_score(z, z, z, z), np.ones((1, 3), dtype=np.float32))  # zero_division=1


def test_normalize_matches_reference_golden():
    from oracle.resnet_unet import normalize
    g = np.load(os.path.join(GOLD, "normalize_reference.npz"))
    out = normalize(torch.from_numpy(g["batch"]), torch.tensor([0.485, 0.456, 0.406]),
                    torch.tensor([0.229, 0.224, 0.225]))
    assert np.array_equal(out.numpy(), g["normed"])


def test_resnet_unet_matches_reference_golden():
    from oracle.resnet_unet import ResNetUNet
    g = np.load(os.path.join(GOLD, "resnet_unet_reference.npz"))
    torch.manual_seed(6210)
    model = ResNetUNet(3, 18).eval()
    sd = model.state_dict()
    for k, s in zip(g["keys"], g["sums"]):
        assert abs(float(sd[str(k)].double().sum()) - float(s)) <= 1e-9 * max(1.0, abs(float(s))), k
    with torch.no_grad():
        y = model(torch.from_numpy(g["x"]))
    assert np.allclose(y.numpy(), g["logits"], rtol=0, atol=1e-5)


def _oracle_unet_like_golden(bilinear=True):
    from oracle.unet import UNet
    torch.manual_seed(6210)
    model = UNet(3, 3, bilinear=bilinear)
    g = torch.Generator().manual_seed(6211)
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.weight.data.uniform_(0.5, 1.5, generator=g)
            m.bias.data.normal_(0, 0.2, generator=g)
            m.running_mean.normal_(0, 0.2, generator=g)
            m.running_var.uniform_(0.5, 1.5, generator=g)
    return model


def test_unet_matches_reference_golden():
    """oracle/unet.py against the reference's own UNet(3, 3, bilinear=True) run by oracle/make_golden.py:
    identical weights from the same seed, eval- and train-mode logits."""
    g = np.load(os.path.join(GOLD, "unet_reference.npz"))
    model = _oracle_unet_like_golden()
    sd = model.state_dict()
    assert [str(k) for k in g["keys"]] == [k for k, v in sd.items() if v.dtype.is_floating_point]
    for k, s in zip(g["keys"], g["sums"]):
        assert abs(float(sd[str(k)].double().sum()) - float(s)) <= 1e-9 * max(1.0, abs(float(s))), k
    x = torch.from_numpy(g["x"])
    model.eval()
    with torch.no_grad():
        assert np.allclose(model(x).numpy(), g["logits_eval"], rtol=0, atol=1e-5)
    model.train()
    with torch.no_grad():
        assert np.allclose(model(x).numpy(), g["logits_train"], rtol=0, atol=1e-5)
    assert sum(p.numel() for p in model.parameters()) == 17267523     # 17.27 M


def test_unet_convtranspose_matches_reference_golden():
    """The bilinear=False variant (ConvTranspose2d(in, in // 2, 2, 2) upsampling, unet_parts.py:269; down4 widens
    to 1024 channels, unet.py:153-163) against the reference's own UNet(3, 3, bilinear=False)."""
    g = np.load(os.path.join(GOLD, "unet_convt_reference.npz"))
    model = _oracle_unet_like_golden(bilinear=False)
    sd = model.state_dict()
    assert [str(k) for k in g["keys"]] == [k for k, v in sd.items() if v.dtype.is_floating_point]
    for k, s in zip(g["keys"], g["sums"]):
        assert abs(float(sd[str(k)].double().sum()) - float(s)) <= 1e-9 * max(1.0, abs(float(s))), k
    x = torch.from_numpy(g["x"])
    model.eval()
    with torch.no_grad():
        assert np.allclose(model(x).numpy(), g["logits_eval"], rtol=0, atol=1e-5)
    model.train()
    with torch.no_grad():
        assert np.allclose(model(x).numpy(), g["logits_train"], rtol=0, atol=1e-5)
    assert sum(p.numel() for p in model.parameters()) == 31043651     # 31.04 M
    assert tuple(sd["up1.up.weight"].shape) == (1024, 512, 2, 2)


def test_unetpp_structure_pins():
    """README torchinfo dump of UnetPlusPlus + mobilenetv3 (MMR_EN:DE_CODER/README.md:149-188):
    the restated decoder formula reproduces all 11 block parameter counts and their sum; the
    resnet18 model has the '~15M parameters' of README.md:54."""
    from oracle.unetpp import UnetPlusPlus, decoder_block_specs, decoder_schedule
    pins = {"x_0_0": 2028544, "x_1_1": 20832, "x_2_2": 8128, "x_3_3": 6976, "x_0_1": 498176, "x_1_2": 10432,
            "x_2_3": 9280, "x_0_2": 138496, "x_1_3": 11584, "x_0_3": 46208, "x_0_4": 6976}
    specs = decoder_block_specs((3, 16, 16, 24, 48, 576))
    got = {k: (i + s) * o * 9 + 2 * o + o * o * 9 + 2 * o for k, (i, s, o) in specs.items()}
    assert got == pins and sum(got.values()) == 2785632
    assert 16 * 10 * 9 + 10 == 1450  # 3x3 head with bias at 10 classes
    m = UnetPlusPlus("resnet18", None, 3, 2)
    assert sum(p.numel() for p in m.parameters()) == 15970594
    assert sum(p.numel() for p in UnetPlusPlus("resnet18", None, 3, 10).parameters()) == 15971754
    order = [b for b, _, _ in decoder_schedule()]
    assert order == ["x_0_0", "x_1_1", "x_2_2", "x_3_3", "x_0_1", "x_1_2", "x_2_3", "x_0_2", "x_1_3", "x_0_3", "x_0_4"]
    sd = m.state_dict()
    for key in ("encoder.conv1.weight", "encoder.layer2.0.downsample.0.weight", "encoder.layer4.1.bn2.running_var",
                "decoder.blocks.x_0_3.conv1.0.weight", "decoder.blocks.x_1_2.conv2.1.num_batches_tracked",
                "segmentation_head.0.bias"):
        assert key in sd, key
    assert tuple(sd["decoder.blocks.x_0_3.conv1.0.weight"].shape) == (32, 320, 3, 3)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 48, 64))


def test_dice_loss_restatement_properties():
    """dice_loss (SU/dice_loss.py:118-159) with kornia's +1e-6 one-hot: hand-computed tiny case,
    error behaviour, and DiceCE at uniform predictions = ln(10) + 0.9 (the reference's
    loss_dict.pkl starts at 2.786 on its imbalanced labels: same order)."""
    from oracle import losses
    logits = torch.tensor([[[[2.0, 0.0]], [[0.0, 2.0]]]])          # N=1, C=2, H=1, W=2
    target = torch.tensor([[[0, 0]]])
    p = torch.softmax(logits, 1)
    y = torch.tensor([[[[1.0, 1.0]], [[0.0, 0.0]]]]) + 1e-6
    inter = (p * y).sum((2, 3))
    card = (p + y).sum((2, 3))
    want = (1 - (2 * inter + 1.0) / (card + 1.0)).mean()
    assert torch.allclose(losses.dice_loss(logits, target), want, atol=1e-7)
    with pytest.raises(TypeError):
        losses.dice_loss([1, 2], target)
    with pytest.raises(ValueError):
        losses.dice_loss(torch.zeros(2, 3, 4), target)
    with pytest.raises(ValueError):
        losses.dice_loss(torch.zeros(1, 2, 3, 3), target)
    g = torch.Generator().manual_seed(0)
    z = torch.randn((2, 10, 32, 32), generator=g) * 0.01
    t = torch.randint(0, 10, (2, 32, 32), generator=g)
    oh = torch.nn.functional.one_hot(t, 10).permute(0, 3, 1, 2).float()
    v = losses.monai_dice_ce(z, oh).item()
    assert abs(v - (np.log(10) + 0.9)) < 0.02
    assert abs(losses.mixed_loss(z, t, -1).item() - np.log(10)) < 0.01


def test_hausdorff_restatement_known_answers():
    """oracle.metrics.hausdorff_distance (skimage.metrics.hausdorff_distance restated with scipy's cKDTree):
    closed-form cases -- a 3-4-5 triangle, nested sets, and skimage's empty-set conventions (0 / inf)."""
    from oracle.metrics import hausdorff_distance
    a, b = np.zeros((8, 9), bool), np.zeros((8, 9), bool)
    assert hausdorff_distance(a, b) == 0.0
    a[1, 2] = True
    assert hausdorff_distance(a, b) == np.inf and hausdorff_distance(b, a) == np.inf
    b[4, 6] = True
    assert hausdorff_distance(a, b) == 5.0
    b[1, 2] = True                      # B now contains A: the far point of B decides
    assert hausdorff_distance(a, b) == 5.0
    a[4, 5] = True
    assert hausdorff_distance(a, b) == 1.0


def test_smp_unet_restatement_structure():
    """oracle.unetpp.Unet (the reference's `--model smp_unet18`, SU/ModelTraining.py:255-262): smp's published size
    for Unet-resnet18 (14.3 M parameters: 14 328 209 at one class), smp's state_dict key scheme, and the product
    model carries the same keys and shapes."""
    import torch
    from oracle.unetpp import Unet
    from mmrseg_b200.models import Unet as Product
    ref = Unet("resnet18", None, 3, 1)
    assert sum(p.numel() for p in ref.parameters()) == 14_328_209
    sd = ref.state_dict()
    assert sd["decoder.blocks.0.conv1.0.weight"].shape == (256, 768, 3, 3)
    assert sd["decoder.blocks.3.conv1.0.weight"].shape == (32, 128, 3, 3)
    assert sd["decoder.blocks.4.conv1.0.weight"].shape == (16, 32, 3, 3)
    prod = Product("resnet18", classes=1).state_dict()
    assert list(prod.keys()) == list(sd.keys()) and all(prod[k].shape == sd[k].shape for k in sd)
    ref.eval()
    with torch.no_grad():
        assert ref(torch.zeros(1, 3, 64, 96)).shape == (1, 1, 64, 96)
